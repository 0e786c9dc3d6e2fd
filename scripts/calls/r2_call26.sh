B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
B10="python bench.py --only c3 --steps 3 --warmup 2 --no-cpu-baseline --no-peaks --no-est"
for v in 1 0 1 0; do
  timeout 600 $B --opt tail_spread=$v > gpurun_out/spread_$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/spread_$v.log") if l.startswith("{")][-1])
print("4M spread $v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
  timeout 600 $B10 --opt tail_spread=$v > gpurun_out/spread10_$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/spread10_$v.log") if l.startswith("{")][-1])
print("10M spread $v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
timeout 900 python scripts/fuzz_parity.py 150 92 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -1 gpurun_out/pytest_gpu.log
