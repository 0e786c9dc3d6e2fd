show() { python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$1.log") if l.startswith("{")][-1])
    print("$1: value", round(d["value"]/1e6,2), "ms", round(d["ms_per_step"],2), "kernel", round(d["roofline"]["kernel_ms_per_step"],2), d["cascade_rows_per_step"])
except Exception as ex:
    print("$1 failed", ex); print(open("gpurun_out/$1.log").read()[-800:])
PY
}
BQ="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in sp1 sp2 sp3 main; do
  if [ $v = main ]; then unset SKNNR_B200_LIB; else export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_$v.so; fi
  timeout 300 $BQ > gpurun_out/q_$v.log 2>&1; show q_$v
done
unset SKNNR_B200_LIB
for st in 3 5 6; do timeout 300 $BQ --tc-seed-stride $st > gpurun_out/q_st$st.log 2>&1; show q_st$st; done
P1="python bench.py --only c3 --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-peaks --no-e2e --no-est"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 2 -c 1 -o gpurun_out/prof_tc_r02f -f $P1 > gpurun_out/ncu_tc.log 2>&1; echo ncu_exit=$?
