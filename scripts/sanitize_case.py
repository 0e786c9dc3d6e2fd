"""CI-sized run of every search / refine kernel for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_case.py

Shapes are small (the tools slow kernels down 10-100x) but take every branch of the cascade:
tensor filter (both stream layouts, seeding on), SIMT engine, exhaustive engine (k <= 32 and k > 32),
refine, Hamming with equal and unequal weights, forest walk, raster front end."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sknnr_b200 import _lib as L
from sknnr_b200._engine import ForestIndex, HammingIndex, KNNIndex

rng = np.random.default_rng(0)
n_ref, d, n_q = 9000, 32, 700        # 71 reference tiles of 128: the seeding pass is on (>= 64 tiles)
R = rng.standard_normal((n_ref, d))
R[500:520] = R[:20]
y = rng.standard_normal((n_ref, 3))
Q = np.vstack([R[rng.integers(0, n_ref, n_q // 2)], rng.standard_normal((n_q - n_q // 2, d))])
ix = KNNIndex(R, None, None, None, y, device=0)
ref = None
for eng, name in ((L.ENGINE_EXACT, "exact"), (L.ENGINE_TENSOR, "tensor"), (L.ENGINE_SIMT, "simt")):
    L.set_option("engine", eng)
    out = ix.query(Q, 7, transformed=True, weights="distance", with_pred=True)
    print(name, ix.stats())
    if ref is None:
        ref = out
    assert np.array_equal(out[1], ref[1]) and np.array_equal(out[0], ref[0]), name
L.set_option("engine", L.ENGINE_AUTO)
a = ix.query(Q, 12, transformed=True)                 # single-stream layout of the tensor engine
b = ix.query(None, 5, exclude_self=True)
c = ix.query(Q[:40], 40, transformed=True)            # k > 32: exhaustive engine, CTA-wide finish
print("aux ok", a[1].shape, b[1].shape, c[1].shape)

T = 100
Rc = rng.integers(0, 30, size=(3000, T)).astype(np.uint16)
Qc = Rc[rng.integers(0, 3000, 400)].copy()
Qc[rng.random(Qc.shape) < 0.3] = 31
for w in (np.full(T, 1.0 / T), rng.random(T)):
    hx = HammingIndex(Rc, w, y[:3000], device=0)
    h = hx.query(Qc, 7, weights="uniform", with_pred=True)
    print("hamming", hx.stats())

from sklearn.ensemble import RandomForestRegressor

Xt = rng.standard_normal((400, 5))
rf = RandomForestRegressor(n_estimators=6, random_state=0, min_samples_leaf=5).fit(Xt, Xt[:, 0])
fx = ForestIndex([t.tree_ for t in rf.estimators_], 5, device=0)
ids = fx.apply(rng.standard_normal((300, 5)))
print("forest", ids.shape)

bands = rng.standard_normal((d, 2000)).astype(np.float32)
bands[3, ::17] = np.nan
r = ix.query_raster(bands, 3, weights="uniform", with_pred=True)
print("raster valid", r[3])
print("SANITIZE_CASE_OK")
