set -x
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "certifies_nearly or tensor_engine_conf" 2>&1 | tail -8
timeout 300 python scripts/est_profile.py 10000000 2>&1 | sed -n '1p;4,6p;8p'
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo bench_exit=$?; python - <<'PY'
import json
line=[l for l in open("gpurun_out/bench_full.log") if l.startswith("{")][-1]
d=json.loads(line)
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], d["e2e_estimator"]["per_call_ms"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
print("c5", d["c5"]["value"], d["c5"]["roofline"]["frac"], d["c5"]["fallback_rows_per_step_rank0"], "c4", d["c4"]["value"], d["c4"]["roofline"]["frac"])
PY
