"""Share of rows the tensor engine cannot certify (they are re-searched by the FP32 engine) over
reference-set sizes and dimensions: the library demotes an index above 5 %."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex
rng = np.random.default_rng(0)
for d in (8, 17, 32, 64):
    row = []
    for n_ref in (300, 2000, 5000, 9000, 20000, 50000):
        R = rng.standard_normal((n_ref, d)); Q = rng.standard_normal((20000, d))
        ix = KNNIndex(R, None, None, None, None)
        L.set_option("engine", L.ENGINE_TENSOR)
        ix.query(Q, 7, transformed=True)
        st = ix.stats()
        row.append(f"{n_ref}:{100.0 * st['n_fallback'] / st['n_queries']:.2f}%")
        L.set_option("engine", L.ENGINE_AUTO)
    print(f"d={d:3d} k=7 ", "  ".join(row), flush=True)
