set -x
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for v in "" _early; do
export SKNNR_B200_LIB=$PWD/sknnr_b200/lib/libsknnr_b200$v.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "certifies_nearly" 2>&1 | tail -3
timeout 300 $B > gpurun_out/bench_v$v.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_v$v.log") if l.startswith("{")][-1])
print("variant '$v' value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
done
