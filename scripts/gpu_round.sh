set -x
timeout 300 python scripts/tc_smoke.py > gpurun_out/tc_smoke.log 2>&1; echo tc_smoke_exit=$?; grep -c OK gpurun_out/tc_smoke.log; grep -v OK gpurun_out/tc_smoke.log | tail -5
B="python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_ns2.log 2>&1; grep -o '"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*\|"fallback_rows_per_step": [0-9]*' gpurun_out/bench_ns2.log | tr '\n' ' '; echo
timeout 300 $B --tc-debug 1 > gpurun_out/bench_dbg.log 2>&1; grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/bench_dbg.log
