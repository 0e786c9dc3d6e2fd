set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/bench_default.log 2>&1; echo bench_exit=$?; grep -o '"value": [0-9.]*\|"kernel_ms_per_step": [0-9.]*\|"ms_per_step": [0-9.]*\|"fallback_rows_per_step": [0-9]*\|"frac": [0-9.]*' gpurun_out/bench_default.log | tr '\n' ' '; echo
