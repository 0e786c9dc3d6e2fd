set -x
timeout 300 python scripts/tc_smoke.py > gpurun_out/tc_smoke.log 2>&1; echo tc_smoke_exit=$?; tail -22 gpurun_out/tc_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 1 --n-queries 4194304 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; echo bench_exit=$?; tail -c 2600 gpurun_out/bench_tc.log | head -c 1800
