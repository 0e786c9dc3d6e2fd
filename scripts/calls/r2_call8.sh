# round 2, call 8: full GPU suite on the default build, self-check build (device assertions), timing
set -x
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -12 gpurun_out/pytest_gpu.log
export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so
timeout 900 python scripts/sanitize_case.py > gpurun_out/selfcheck_case.log 2>&1; echo selfcheck_case_exit=$?; tail -2 gpurun_out/selfcheck_case.log
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/selfcheck_pytest.log 2>&1; echo selfcheck_pytest_exit=$?; tail -3 gpurun_out/selfcheck_pytest.log
timeout 900 python scripts/fuzz_parity.py 250 41 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -2 gpurun_out/selfcheck_fuzz.log
grep -c "SK_CHECK failed" gpurun_out/selfcheck_*.log
unset SKNNR_B200_LIB
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304"
G='"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
timeout 300 $B > gpurun_out/bench_default.log 2>&1; echo "exit=$?"; tail -c 3000 gpurun_out/bench_default.log | grep -o "$G" | tr '\n' ' '; echo
timeout 300 $B --dim 64 > gpurun_out/bench_default_d64.log 2>&1; echo "exit=$?"; tail -c 3000 gpurun_out/bench_default_d64.log | grep -o "$G" | tr '\n' ' '; echo
