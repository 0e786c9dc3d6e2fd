import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sknnr_oracle as orc
from sknnr_b200._engine import KNNIndex
np.set_printoptions(linewidth=200, precision=4)
rng = np.random.default_rng(33)
n_ref = 301
R = rng.standard_normal((n_ref, 6)); y = rng.standard_normal((n_ref, 3))
Q = rng.standard_normal((5, 6))
st = orc.FittedState("euclidean", fit_Z=R, y=y)
ix = KNNIndex(R, None, None, None, y)
for k in (33, 40):
    d_o, i_o = orc.kneighbors(st, Q, k=k, transformed=True)
    d_g, i_g, p_g = ix.query(Q, k, transformed=True, weights="distance", with_pred=True)
    print("k", k, "stats", ix.stats())
    print("ours d", d_g[0]); print("orc  d", d_o[0]); print("ours i", i_g[0]); print("orc  i", i_o[0])
    print("nondet:"); d2, i2, _ = ix.query(Q, k, transformed=True, deterministic=False); print(d2[0]); print(i2[0])
