set -x
B="python bench.py --steps 2 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
for dbg in 0 16 17; do timeout 300 $B --tc-debug $dbg > gpurun_out/bench_d$dbg.log 2>&1; echo dbg=$dbg; grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/bench_d$dbg.log; done
