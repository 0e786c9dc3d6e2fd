"""Quick check of the tensor-core engine against the CPU oracle (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sknnr_oracle as orc
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex

def run(n_ref, n_q, d, k, engine):
    rng = np.random.default_rng(0)
    R = rng.standard_normal((n_ref, d)); Q = rng.standard_normal((n_q, d))
    y = rng.standard_normal((n_ref, 3))
    st = orc.FittedState("euclidean", fit_Z=R, y=y)
    ix = KNNIndex(R, y=y)
    L.set_option("engine", engine)
    t0 = time.time()
    dg, ig, pg = ix.query(Q, k, transformed=True, weights="distance", with_pred=True)
    dt = time.time() - t0
    stt = ix.stats()
    do, io = orc.kneighbors(st, Q, k=k, transformed=True)
    nbad = orc.assert_tie_aware_equal(dg, ig, do, io, rtol=1e-5, atol=1e-7)
    print(f"engine={engine} n_ref={n_ref} n_q={n_q} d={d} k={k}: OK rows_differ={nbad} "
          f"fallback={stt['n_fallback']} used_engine={stt['engine']} launches={stt['kernel_launches']} {dt*1e3:.1f} ms", flush=True)

if __name__ == "__main__":
    for args in [(300, 200, 8, 3), (1000, 700, 32, 7), (5000, 3000, 32, 7), (20000, 4000, 64, 7), (777, 513, 17, 5), (4096, 1024, 40, 12),
                 (50000, 3000, 32, 7), (9000, 2000, 24, 15)]:
        for engine in (2, 1):
            run(*args, engine)
    for ns in (1, 2):
        L.set_option("tc_streams", ns)
        for args in [(300, 200, 8, 3), (5000, 3000, 32, 7), (50000, 3000, 32, 7), (20000, 1500, 24, 7), (9000, 2000, 24, 15)]:
            run(*args, 2)
    L.set_option("tc_streams", 0)
    for stride in (0, 2, 8):
        L.set_option("tc_seed_stride", stride)
        run(50000, 3000, 32, 7, 2)
    L.set_option("tc_seed_stride", 4)
    L.set_option("engine", 0)
    print("tc_smoke done")
