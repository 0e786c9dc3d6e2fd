python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 200 103 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/selfcheck_pytest.log 2>&1; echo selfcheck_exit=$?; tail -1 gpurun_out/selfcheck_pytest.log; grep -c "SK_CHECK failed" gpurun_out/selfcheck_pytest.log
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo bench_exit=$?; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_full.log") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], d["e2e_estimator"]["per_call_ms"], "kernel", d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["frac_tf32_basis"], d["cascade_rows_per_step"])
print("c5", d["c5"]["value"], d["c5"]["roofline"]["frac"], d["c5"]["cascade_rows_per_step_rank0"], "c4", d["c4"]["value"], d["c4"]["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["value"])
PY
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo ref_exit=$?; tail -1 gpurun_out/bench_ref.log | cut -c1-600
