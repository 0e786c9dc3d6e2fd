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
timeout 600 $BE > gpurun_out/e_base.log 2>&1; show e_base
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 $BE > gpurun_out/e_conn32.log 2>&1; show e_conn32
timeout 600 $BE --host-slots 3 > gpurun_out/e_s3.log 2>&1; show e_s3
timeout 600 $BE --host-slots 4 > gpurun_out/e_s4.log 2>&1; show e_s4
timeout 600 $BE --host-slots 4 --chunk-rows 2097152 > gpurun_out/e_s4c2.log 2>&1; show e_s4c2
timeout 600 $BE --chunk-rows 524288 --no-est > gpurun_out/e_c512.log 2>&1; show e_c512
timeout 900 python scripts/fuzz_parity.py 150 101 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python scripts/fuzz_parity.py 100 102 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -1 gpurun_out/selfcheck_fuzz.log; grep -c "SK_CHECK failed" gpurun_out/selfcheck_fuzz.log
