showc4() { python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/$1.log") if l.startswith("{")][-1])
    c=d.get("c4") or d
    print("$1: c4 value", c["value"], "ms", c["ms_per_step"], "frac", c["roofline"]["frac"], "kernel_ms", c["roofline"]["kernel_ms_per_step"])
except Exception as ex:
    print("$1 failed", ex); print(open("gpurun_out/$1.log").read()[-1200:])
PY
}
B="python bench.py --only c4 --no-cpu-baseline --no-peaks --steps 2 --warmup 1"
export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_hw16.so
timeout 600 $B > gpurun_out/c4_hw16.log 2>&1; showc4 c4_hw16
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_estimators.py -m gpu -q -k "hamming or rfnn or gbnn or forest or c4" > gpurun_out/pytest_ham.log 2>&1; echo pytest_exit=$?; tail -2 gpurun_out/pytest_ham.log
timeout 300 python scripts/hamming_bench.py weighted 2>&1 | tail -3
unset SKNNR_B200_LIB
timeout 600 $B > gpurun_out/c4_main.log 2>&1; showc4 c4_main
timeout 300 python scripts/hamming_bench.py weighted 2>&1 | tail -3
