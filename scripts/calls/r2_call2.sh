# round 2, call 2: queue/drain kernel variants, timing experiments, ncu of the q kernel
set -x
L=$PWD/sknnr_b200/lib
B="python bench.py --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-e2e"
G='"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
for v in q t3 j10 j16; do
  SKNNR_B200_LIB=$L/libsknnr_b200_$v.so timeout 300 $B > gpurun_out/bench_$v.log 2>&1; echo "$v exit=$?"; tail -c 1500 gpurun_out/bench_$v.log | grep -o "$G"
done
export SKNNR_B200_LIB=$L/libsknnr_b200_q.so
for dbg in 1 32; do
  timeout 300 $B --tc-debug $dbg > gpurun_out/bench_q_dbg$dbg.log 2>&1; echo "dbg$dbg exit=$?"; tail -c 1500 gpurun_out/bench_q_dbg$dbg.log | grep -o "$G"
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo pytest_q_exit=$?; tail -5 gpurun_out/pytest_q.log
P="python bench.py --steps 1 --warmup 1 --n-queries 1048576 --no-cpu-baseline --no-e2e"
timeout 300 $P > gpurun_out/plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_tc -s 1 -c 1 -o gpurun_out/prof_tcq $P > gpurun_out/ncu_tcq.log 2>&1; echo ncu_exit=$?
