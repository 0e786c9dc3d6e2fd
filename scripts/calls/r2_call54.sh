timeout 45 python bench.py --only c3 --steps 3 --warmup 3 --no-cpu-baseline --no-peaks --no-est > gpurun_out/final_c3.log 2>&1; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/final_c3.log") if l.startswith("{")][-1])
print("value", round(d["value"]/1e6,2), "e2e", round(d["e2e"]["value"]/1e6,2), "kernel", round(d["roofline"]["kernel_ms_per_step"],2), "launches", d["gpu_launches"])
PY
