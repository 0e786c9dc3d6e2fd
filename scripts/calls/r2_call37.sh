timeout 120 python scripts/pcie_bw.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "second_pass or joint_threshold or host or pageable" > gpurun_out/pytest_q.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_q.log
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
timeout 600 $BE > gpurun_out/cs1.log 2>&1; show cs1
timeout 600 $BE --opt copy_streams=0 > gpurun_out/cs0.log 2>&1; show cs0
timeout 600 $BE --opt copy_streams=1 --host-slots 4 > gpurun_out/cs1_s4.log 2>&1; show cs1_s4
BQ="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in ov1 ov2 main; do
  if [ $v = main ]; then unset SKNNR_B200_LIB; else export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_$v.so; fi
  timeout 600 $BQ > gpurun_out/q_$v.log 2>&1; show q_$v
done
unset SKNNR_B200_LIB
