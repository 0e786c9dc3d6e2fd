set -x
L=$PWD/sknnr_b200/lib
for v in "" _vote; do echo "== variant '$v'"; SKNNR_B200_LIB=$L/libsknnr_b200$v.so timeout 300 python scripts/fallback_sweep.py 2>&1 | tail -4; done
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in "" _vote; do
SKNNR_B200_LIB=$L/libsknnr_b200$v.so timeout 300 $B > gpurun_out/bench_x$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_x$v.log") if l.startswith("{")][-1])
print("variant '$v' value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
done
