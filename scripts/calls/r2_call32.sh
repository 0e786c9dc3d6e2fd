timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 250 96 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/selfcheck_pytest.log 2>&1; echo selfcheck_exit=$?; tail -2 gpurun_out/selfcheck_pytest.log; grep -c "SK_CHECK failed" gpurun_out/selfcheck_pytest.log
timeout 300 python scripts/fallback_sweep.py 2>&1 | tail -12
