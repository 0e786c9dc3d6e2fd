"""Which true neighbours does the tensor engine miss? (run on the GPU box)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sknnr_oracle as orc
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex

def run(n_ref, n_q, d, k):
    rng = np.random.default_rng(0)
    R = rng.standard_normal((n_ref, d)); Q = rng.standard_normal((n_q, d))
    st = orc.FittedState("euclidean", fit_Z=R, y=None)
    ix = KNNIndex(R)
    L.set_option("engine", 2)
    dg, ig, _ = ix.query(Q, k, transformed=True)
    do, io = orc.kneighbors(st, Q, k=k, transformed=True)
    bad = 0
    for q in range(n_q):
        miss = sorted(set(io[q]) - set(ig[q]))
        if miss:
            bad += 1
            if bad <= 40:
                ranks = [int(np.where(io[q] == m)[0][0]) for m in miss]
                print(f"q={q} (lane {q%32}, warp {q//32}) missed refs {miss} (tile {[m//128 for m in miss]}, chunk {[(m%128)//32 for m in miss]}, col {[m%32 for m in miss]}) ranks {ranks}; got {list(ig[q])}")
    print(f"n_ref={n_ref} n_q={n_q} d={d} k={k}: rows with misses {bad}, stats {ix.stats()}", flush=True)

for args in [(300, 200, 8, 3), (128, 64, 8, 3), (32, 32, 8, 3), (64, 32, 8, 3), (1000, 700, 32, 7)]:
    run(*args)
