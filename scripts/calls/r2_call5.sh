# round 2, call 5: large-k tests, new bench (C3 + C5 + C4 + estimator e2e), seeding stride sweep
set -x
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "large_k or error_behaviour" 2>&1 | tail -5
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo bench_exit=$?; tail -5 gpurun_out/bench_full.err | cut -c1-900; tail -c 12000 gpurun_out/bench_full.log
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304"
G='"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
for ss in 0 16 6; do
  timeout 300 $B --tc-seed-stride $ss > gpurun_out/bench_ss$ss.log 2>&1; echo "ss$ss exit=$?"; tail -c 3000 gpurun_out/bench_ss$ss.log | grep -o "$G" | tr '\n' ' '; echo
done
