BE="python bench.py --only c3 --steps 3 --warmup 3 --no-cpu-baseline --no-peaks"
show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$1.log") if l.startswith("{")][-1])
    e=d.get("e2e") or {}; es=d.get("e2e_estimator") or {}
    print("$1: value", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],2), "kernel", round(d["roofline"]["kernel_ms_per_step"],2), "e2e", round((e.get("value") or 0)/1e6,2), "est", round((es.get("value") or 0)/1e6,2), es.get("per_call_ms"), d["cascade_rows_per_step"])
except Exception as ex:
    print("$1 failed", ex); print(open("gpurun_out/$1.log").read()[-800:])
PY
}
timeout 600 $BE --chunk-rows 131072 > gpurun_out/e_c128.log 2>&1; show e_c128
timeout 600 $BE --chunk-rows 262144 > gpurun_out/e_c256.log 2>&1; show e_c256
timeout 600 $BE --chunk-rows 393216 > gpurun_out/e_c384.log 2>&1; show e_c384
timeout 600 $BE --chunk-rows 524288 --opt stage_rows=262144 > gpurun_out/e_c512s256.log 2>&1; show e_c512s256
timeout 600 $BE --chunk-rows 262144 --host-slots 4 > gpurun_out/e_c256s4.log 2>&1; show e_c256s4
