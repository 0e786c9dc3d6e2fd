timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "second_pass" > gpurun_out/pytest_2p.log 2>&1; echo pytest2p_exit=$?; tail -15 gpurun_out/pytest_2p.log
B="python bench.py --only c3 --steps 4 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
B10="python bench.py --only c3 --steps 3 --warmup 2 --no-cpu-baseline --no-peaks --no-est"
for v in 1 0 1 0; do
  timeout 600 $B --opt tc_retry=$v > gpurun_out/retry_$v.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/retry_$v.log") if l.startswith("{")][-1])
    print("4M retry $v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], d["cascade_rows_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/retry_$v.log").read()[-1500:])
PY
  timeout 600 $B10 --opt tc_retry=$v > gpurun_out/retry10_$v.log 2>&1; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/retry10_$v.log") if l.startswith("{")][-1])
    print("10M retry $v: value", d["value"], "ms", d["ms_per_step"], "kernel", d["roofline"]["kernel_ms_per_step"], "e2e", d["e2e"]["value"], d["cascade_rows_per_step"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/retry10_$v.log").read()[-1500:])
PY
done
timeout 900 python scripts/fuzz_parity.py 200 94 > gpurun_out/fuzz.log 2>&1; echo fuzz_exit=$?; tail -2 gpurun_out/fuzz.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_gpu.log
SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200_checks.so timeout 900 python scripts/fuzz_parity.py 100 95 > gpurun_out/selfcheck_fuzz.log 2>&1; echo selfcheck_fuzz_exit=$?; tail -1 gpurun_out/selfcheck_fuzz.log
