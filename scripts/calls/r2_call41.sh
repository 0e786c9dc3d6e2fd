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
SKNNR_B200_TRACE=1 timeout 600 $BE --no-est --steps 1 > gpurun_out/tr_p1.log 2> gpurun_out/tr_p1.err; grep -A40 "sknnr trace" gpurun_out/tr_p1.err | tail -15
timeout 600 $BE > gpurun_out/p1.log 2>&1; show p1
timeout 600 $BE --opt host_pipeline=0 > gpurun_out/p0.log 2>&1; show p0
timeout 600 $BE --chunk-rows 524288 > gpurun_out/p1_c512.log 2>&1; show p1_c512
timeout 600 $BE --host-slots 3 > gpurun_out/p1_s3.log 2>&1; show p1_s3
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_estimators.py -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_q.log
