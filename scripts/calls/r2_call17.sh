set -x
timeout 120 python - <<'PY'
import numpy as np, sys
sys.path.insert(0,'.')
from oracle import sknnr_oracle as orc
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex
rng=np.random.default_rng(0)
R=rng.standard_normal((9000,32)); y=rng.standard_normal((9000,3)); Q=rng.standard_normal((3000,32))
ix=KNNIndex(R,None,None,None,y)
L.set_option("engine", L.ENGINE_EXACT); ref=ix.query(Q,7,transformed=True)
L.set_option("engine", L.ENGINE_AUTO)
for lay in (2,4):
    L.set_option("tc_layout", lay)
    out=ix.query(Q,7,transformed=True); st=ix.stats()
    print("layout",lay,"equal idx",np.array_equal(out[1],ref[1]),"equal d",np.array_equal(out[0],ref[0]),st)
PY
echo quick_exit=$?
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -5
timeout 600 python scripts/fuzz_parity.py 150 61 > gpurun_out/fuzz4.log 2>&1; echo fuzz_exit=$?; tail -2 gpurun_out/fuzz4.log
B="python bench.py --only c3 --steps 3 --warmup 2 --n-queries 4194304 --no-cpu-baseline --no-peaks --no-e2e --no-est"
for lay in 4 2; do
timeout 300 $B --tc-layout $lay > gpurun_out/bench_l$lay.log 2>&1; echo "layout $lay exit=$?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_l$lay.log") if l.startswith("{")][-1])
    print("layout $lay value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
except Exception as e: print("fail", e); print(open("gpurun_out/bench_l$lay.log").read()[-1500:])
PY
done
timeout 300 $B --tc-layout 4 --dim 64 > gpurun_out/bench_l4_d64.log 2>&1; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_l4_d64.log") if l.startswith("{")][-1])
print("d64 layout 4 value", d["value"], "kernel", d["roofline"]["kernel_ms_per_step"], "fb", d["fallback_rows_per_step"])
PY
