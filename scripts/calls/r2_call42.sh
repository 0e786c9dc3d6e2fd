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
timeout 600 $BE > gpurun_out/n_base.log 2>&1; show n_base
timeout 600 $BE --opt tail_priority=0 > gpurun_out/n_tp0.log 2>&1; show n_tp0
timeout 600 $BE --opt host_nt=0 > gpurun_out/n_nt0.log 2>&1; show n_nt0
timeout 600 $BE --opt host_threads=16 > gpurun_out/n_t16.log 2>&1; show n_t16
timeout 600 $BE --opt host_threads=12 > gpurun_out/n_t12.log 2>&1; show n_t12
timeout 600 $BE --opt stage_rows=1048576 > gpurun_out/n_st1m.log 2>&1; show n_st1m
SKNNR_B200_TRACE=1 timeout 600 $BE --no-est --steps 1 > gpurun_out/tr_p2.log 2> gpurun_out/tr_p2.err; grep -A40 "sknnr trace" gpurun_out/tr_p2.err | tail -14
