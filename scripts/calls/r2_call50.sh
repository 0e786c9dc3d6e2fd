timeout 900 python bench.py > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err; echo bench_exit=$?; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_full.log") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "est", d["e2e_estimator"]["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], d["cascade_rows_per_step"], d["clocks"])
print("c5", d["c5"]["value"], "c4", d["c4"]["value"], "cpu", d["cpu_baseline"]["value"])
PY
