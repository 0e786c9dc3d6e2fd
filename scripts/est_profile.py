"""Where the time of est.predict(X) goes on a pageable 10M x 32 float64 array (round-2 item: the
drop-in call must come close to the pinned-buffer C-ABI call)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sknnr_b200 import EuclideanKNNRegressor, _lib as L
from sknnr_b200._engine import KNNIndex, pinned_empty
from sklearn.utils.validation import validate_data

n_q = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
R = np.random.default_rng(0).standard_normal((50_000, 32)); y = np.random.default_rng(1).standard_normal((50_000, 8))
est = EuclideanKNNRegressor(n_neighbors=7, weights="distance").fit(R, y)
X = np.random.default_rng(2).standard_normal((n_q, 32))
def t(label, f, n=4):
    out = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); out.append(time.perf_counter() - t0)
    print(f"{label:45s}", " ".join(f"{x*1e3:8.1f}" for x in out), "ms", flush=True)
    return r
t("est.predict(X) pageable", lambda: est.predict(X))
t("transformer._validate_query(finite=False)", lambda: est.transformer_._validate_query(X, finite=False))
t("transformer._validate_query(finite=True)", lambda: est.transformer_._validate_query(X))
ix = est.regressor_._get_index()
t("ix.query pageable in, pooled pinned out", lambda: ix.query(X, 7, weights="distance", with_pred=True, return_distance=False, return_index=False, check_finite=True))
outp = np.empty((n_q, 8)); 
t("ix.query pageable in, pageable out (given)", lambda: ix.query(X, 7, weights="distance", with_pred=True, return_distance=False, return_index=False, out=(None, None, outp)))
Xp = pinned_empty((n_q, 32)); Xp[:] = X
outpp = pinned_empty((n_q, 8))
t("ix.query pinned in, pinned out (given)", lambda: ix.query(Xp, 7, weights="distance", with_pred=True, return_distance=False, return_index=False, out=(None, None, outpp)))
t("np.empty((n_q, 8)) + first touch", lambda: np.empty((n_q, 8)).fill(0.0))
for ht in (4, 12):
    pass
print("stats", ix.stats())
