set -x
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
