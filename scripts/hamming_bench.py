"""C4-shaped timing of the Hamming (RFNN) path: n_ref plots x T trees, k=7 (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sknnr_b200 import _lib as L
from sknnr_b200._engine import HammingIndex

def main(n_ref=20000, n_q=500000, T=500, k=7):
    rng = np.random.default_rng(0)
    # node ids of ~60-leaf trees; queries are perturbed copies of plots so neighbours are meaningful
    Rc = rng.integers(0, 60, size=(n_ref, T)).astype(np.uint16)
    Qc = Rc[rng.integers(0, n_ref, size=n_q)].copy()
    flip = rng.random(Qc.shape) < 0.5
    Qc[flip] = rng.integers(0, 60, size=int(flip.sum())).astype(np.uint16)
    w = np.full(T, 1.0 / T)
    L.set_option("timing", 1)
    hx = HammingIndex(Rc, w, device=0)
    hx.query(Qc[:50000], k)
    t0 = time.perf_counter()
    d, i, _ = hx.query(Qc, k)
    dt = time.perf_counter() - t0
    st = hx.stats()
    print(f"hamming n_ref={n_ref} T={T} n_q={n_q} k={k}: e2e {n_q/dt/1e6:.3f} M queries/s ({dt*1e3:.1f} ms), "
          f"search kernel {st['search_ms']:.1f} ms = {n_q*n_ref*T/st['search_ms']/1e9:.1f} T id-compares/s, stats {st}")

if __name__ == "__main__":
    main()
