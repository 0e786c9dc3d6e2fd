set -x
L=$PWD/sknnr_b200/lib
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est --tc-layout 4"
for v in "" _nomc _nodb _d8; do
SKNNR_B200_LIB=$L/libsknnr_b200$v.so timeout 300 $B > gpurun_out/bench_t4$v.log 2>&1; echo "variant '$v' exit=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_t4$v.log") if l.startswith("{")][-1])
    print("variant '$v' value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
except Exception as e: print("fail", e)
PY
done
P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-peaks --no-e2e --no-est --tc-layout 4"
timeout 300 $P1 > gpurun_out/plain1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc4 -s 1 -c 1 -o gpurun_out/prof_tc4 $P1 > gpurun_out/ncu_tc4.log 2>&1; echo ncu_exit=$?
