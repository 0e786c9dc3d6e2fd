# round 2, call 7: compute-sanitizer memcheck on the CI-sized case + quick timing of the default build
set -x
timeout 300 python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; echo plain_exit=$?; tail -3 gpurun_out/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck.log python scripts/sanitize_case.py > gpurun_out/sanitize_memcheck.out 2>&1; echo memcheck_exit=$?; tail -3 gpurun_out/sanitize_memcheck.out; tail -12 gpurun_out/sanitizer_memcheck.log
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304"
G='"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
timeout 300 $B > gpurun_out/bench_default.log 2>&1; echo "exit=$?"; tail -c 3000 gpurun_out/bench_default.log | grep -o "$G" | tr '\n' ' '; echo
timeout 300 $B --dim 64 > gpurun_out/bench_default_d64.log 2>&1; echo "exit=$?"; tail -c 3000 gpurun_out/bench_default_d64.log | grep -o "$G" | tr '\n' ' '; echo
