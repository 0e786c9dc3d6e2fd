set -x
timeout 900 python bench.py --no-cpu-baseline --steps 2 > gpurun_out/bench_default.log 2>&1; echo bench_exit=$?; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_default.log | tr '\n' ' '; echo
