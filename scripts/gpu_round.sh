set -x
B="python bench.py --steps 2 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
for s in 3 4 6 8; do echo stride=$s; timeout 300 $B --tc-seed-stride $s > gpurun_out/bench_s$s.log 2>&1; grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/bench_s$s.log; done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
