set -x
timeout 180 python scripts/tc_smoke.py > gpurun_out/tc_smoke.log 2>&1; echo tc_smoke_exit=$?; tail -3 gpurun_out/tc_smoke.log
timeout 600 python bench.py --steps 2 --warmup 1 --n-queries 4194304 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; echo bench_exit=$?; tail -c 2400 gpurun_out/bench_tc.log | head -c 1600
timeout 300 python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e > gpurun_out/plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tc3 python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo ncu_full_exit=$?
