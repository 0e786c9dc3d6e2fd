# round 2, call 3: scheduled (CTA-synchronous) drains
set -x
L=$PWD/sknnr_b200/lib
B="python bench.py --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-e2e"
G='"value": [0-9.]*\|kernel_ms_per_step": [0-9.]*\|fallback_rows_per_step": [0-9]*'
for v in d4 d6 d8 d12; do
  SKNNR_B200_LIB=$L/libsknnr_b200_$v.so timeout 300 $B > gpurun_out/bench_$v.log 2>&1; echo "$v exit=$?"; tail -c 1500 gpurun_out/bench_$v.log | grep -o "$G"
done
export SKNNR_B200_LIB=$L/libsknnr_b200_d6.so
timeout 300 $B --dim 64 > gpurun_out/bench_d6_d64.log 2>&1; echo "d64 exit=$?"; tail -c 1500 gpurun_out/bench_d6_d64.log | grep -o "$G"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d6.log 2>&1; echo pytest_exit=$?; tail -5 gpurun_out/pytest_d6.log
timeout 600 python scripts/fuzz_parity.py 150 23 > gpurun_out/fuzz_d6.log 2>&1; echo fuzz_exit=$?; tail -3 gpurun_out/fuzz_d6.log
