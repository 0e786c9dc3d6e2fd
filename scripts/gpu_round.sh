set -x
B="python bench.py --steps 2 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
timeout 300 $B > gpurun_out/bench_a.log 2>&1; grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/bench_a.log
timeout 600 python -m pytest tests -m gpu -x -q -k "tensor or synthetic_euclid or golden_kneighbors or mahalanobis" 2>&1 | tail -3
