set -x
timeout 300 python scripts/tc_smoke.py > gpurun_out/tc_smoke.log 2>&1; echo tc_smoke_exit=$?; grep -c OK gpurun_out/tc_smoke.log; grep -v OK gpurun_out/tc_smoke.log | tail -5
B="python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
for m in 2 1; do timeout 300 $B --tc-streams $m > gpurun_out/bench_ns$m.log 2>&1; echo exit=$?; grep -o '"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*\|"fallback_rows_per_step": [0-9]*' gpurun_out/bench_ns$m.log | tr '\n' ' '; echo; done
for s in 2 8; do timeout 300 $B --tc-seed-stride $s > gpurun_out/bench_seed$s.log 2>&1; echo exit=$?; grep -o '"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_seed$s.log | tr '\n' ' '; echo; done
