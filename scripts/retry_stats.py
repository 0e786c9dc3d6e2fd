"""Debug build only (-DSK_RETRY_STATS): why do first-pass rows fail?"""
import ctypes as C
import numpy as np
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex

lib = L.load()
rng = np.random.default_rng(31)
for k in (5, 7):
    R = rng.standard_normal((50_000, 32)); y = rng.standard_normal((50_000, 3)); Q = rng.standard_normal((400_000, 32))
    ix = KNNIndex(R, None, None, None, y)
    ix.query(Q, k, transformed=True, weights="distance", with_pred=True)
    out = (C.c_ulonglong * 4)()
    lib.sk_retry_stats(out)
    print("k", k, "cascade", ix.cascade_counts(), "retry_threshold calls", out[0], "kth=inf", out[1], "qn over limit", out[2])
