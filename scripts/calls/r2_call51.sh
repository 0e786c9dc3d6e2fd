BE="python bench.py --only c3 --steps 3 --warmup 3 --no-cpu-baseline --no-peaks --no-est"
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$1.log") if l.startswith("{")][-1])
    e=d.get("e2e") or {}
    print("$1: value", round(d["value"]/1e6,2), "e2e", round((e.get("value") or 0)/1e6,2))
except Exception as ex:
    print("$1 failed", ex); print(open("gpurun_out/$1.log").read()[-800:])
PY
}
timeout 300 $BE --chunk-rows 2097152 > gpurun_out/h_c2m.log 2>&1; show h_c2m
timeout 300 $BE --chunk-rows 1572864 > gpurun_out/h_c15m.log 2>&1; show h_c15m
