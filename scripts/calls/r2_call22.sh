set -x
B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in main allslots main allslots; do
  if [ $v = main ]; then unset SKNNR_B200_LIB; else export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_$v.so; fi
  timeout 600 $B > gpurun_out/ab_$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ab_$v.log") if l.startswith("{")][-1])
print("$v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"])
PY
done
unset SKNNR_B200_LIB
timeout 1800 python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 200 81 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -2 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python scripts/fuzz_parity.py 100 82 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -1 gpurun_out/selfcheck_fuzz.log
