set -x
timeout 300 python scripts/fallback_sweep.py 2>&1 | tail -4
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 250 71 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -2 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/selfcheck_pytest.log 2>&1; echo selfcheck_exit=$?; tail -2 gpurun_out/selfcheck_pytest.log; grep -c "SK_CHECK failed" gpurun_out/selfcheck_pytest.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python scripts/fuzz_parity.py 150 72 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -1 gpurun_out/selfcheck_fuzz.log
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo bench_exit=$?; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_full.log") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], d["e2e_estimator"]["per_call_ms"], "kernel", d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"], "fb", d["fallback_rows_per_step"])
print("c5", d["c5"]["value"], d["c5"]["roofline"]["frac"], d["c5"]["fallback_rows_per_step_rank0"], "c4", d["c4"]["value"], d["c4"]["roofline"]["frac"])
PY
P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $P1 > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tc_r02 $P1 > gpurun_out/ncu_tc.log 2>&1; echo ncu_exit=$?
P4="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 300 $P4 > gpurun_out/plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02.csv $P4 > gpurun_out/ncu_list.log 2>&1; echo ncu_list_exit=$?
