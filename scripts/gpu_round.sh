set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/tc_smoke.py > gpurun_out/tc_smoke.log 2>&1; echo tc_smoke_exit=$?; grep -c OK gpurun_out/tc_smoke.log
timeout 900 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/bench_default.log 2>&1; echo bench_exit=$?; grep -o '"value": [0-9.]*\|"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*\|"fallback_rows_per_step": [0-9]*' gpurun_out/bench_default.log | tr '\n' ' '; echo
B4="python bench.py --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-e2e"
timeout 300 $B4 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_tc9.csv $B4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
