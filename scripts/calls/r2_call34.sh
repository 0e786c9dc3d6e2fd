B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in main j9 j10 j11 j13 main; do
  if [ $v = main ]; then unset SKNNR_B200_LIB; else export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_$v.so; fi
  timeout 600 $B > gpurun_out/j_$v.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/j_$v.log") if l.startswith("{")][-1])
    print("$v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], d["cascade_rows_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/j_$v.log").read()[-800:])
PY
done
