B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for s in 2 3 4 3 2; do
  timeout 600 $B --tc-seed-stride $s > gpurun_out/seed_$s.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/seed_$s.log") if l.startswith("{")][-1])
print("stride $s: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
done
