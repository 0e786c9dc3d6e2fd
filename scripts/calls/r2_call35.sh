B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
run() { # name, lib, extra
  if [ "$2" = main ]; then unset SKNNR_B200_LIB; else export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_$2.so; fi
  timeout 600 $B $3 > gpurun_out/x_$1.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/x_$1.log") if l.startswith("{")][-1])
    print("$1: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], d["cascade_rows_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/x_$1.log").read()[-800:])
PY
  unset SKNNR_B200_LIB
}
run j10 main ""
run j12_bypass j12 ""
run j12_nobypass j12 "--opt simt_min_rows=0"
run j10b main ""
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/fuzz_parity.py 250 97 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -1 gpurun_out/fuzz.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python scripts/fuzz_parity.py 120 98 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -1 gpurun_out/selfcheck_fuzz.log
timeout 300 python scripts/fallback_sweep.py 2>&1 | tail -6
