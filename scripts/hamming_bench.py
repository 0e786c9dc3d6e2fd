"""C4-shaped timing of the Hamming (RFNN) path: n_ref plots x T trees, k=7 (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sknnr_b200 import _lib as L
from sknnr_b200._engine import HammingIndex

def main(n_ref=20000, n_q=500000, T=500, k=7, weighted=False):
    rng = np.random.default_rng(0)
    # node ids of ~60-leaf trees; queries are perturbed copies of plots so neighbours are meaningful
    Rc = rng.integers(0, 60, size=(n_ref, T)).astype(np.uint16)
    Qc = Rc[rng.integers(0, n_ref, size=n_q)].copy()
    flip = rng.random(Qc.shape) < 0.5
    Qc[flip] = rng.integers(0, 60, size=int(flip.sum())).astype(np.uint16)
    w = np.full(T, 1.0 / T)
    if weighted:   # boosting-like weights: geometric decay per 100-tree model, GBNN (scope row f3)
        w = np.tile(0.97 ** np.arange(100), (T + 99) // 100)[:T] * (1 + 0.1 * rng.random(T))
        w /= w.sum()
    L.set_option("timing", 1)
    hx = HammingIndex(Rc, w, device=0)
    hx.query(Qc[:50000], k)
    t0 = time.perf_counter()
    d, i, _ = hx.query(Qc, k)
    dt = time.perf_counter() - t0
    st = hx.stats()
    print(f"hamming{' (unequal weights)' if weighted else ''} n_ref={n_ref} T={T} n_q={n_q} k={k}: e2e {n_q/dt/1e6:.3f} M queries/s ({dt*1e3:.1f} ms), "
          f"search kernel {st['search_ms']:.1f} ms = {n_q*n_ref*T/st['search_ms']/1e9:.1f} T id-compares/s, stats {st}")

def forest_main(n_ref=20000, n_q=1000000, d=16, n_targets=10, n_estimators=50, k=7):
    """BASELINE.md C4 recipe at reduced query count: RFNNRegressor(n_estimators=50) on 10 targets
    (T = 500 trees), raw feature rows in, neighbours out: forest walk + Hamming search fused."""
    import sknnr_b200 as S
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n_ref, d))
    y = X[:, :n_targets] * 2 + np.random.default_rng(1).standard_normal((n_ref, n_targets))
    Q = np.random.default_rng(2).standard_normal((n_q, d))
    t0 = time.perf_counter()
    est = S.RFNNRegressor(n_estimators=n_estimators, n_neighbors=k, random_state=0, n_jobs=-1).fit(X, y)
    print(f"fit (scikit-learn forest training + device self-query): {time.perf_counter()-t0:.1f} s", flush=True)
    tr = est.transformer_
    tr.transform(Q[:20000])
    t0 = time.perf_counter(); ids = tr.transform(Q[:200000]); dt = time.perf_counter() - t0
    print(f"device forest walk (transform, ids to host): {200000/dt/1e6:.2f} M rows/s x {ids.shape[1]} trees")
    t0 = time.perf_counter(); want = np.hstack([e.apply(Q[:20000]) for e in tr.estimators_]); dt_cpu = time.perf_counter() - t0
    print(f"scikit-learn est.apply on the host: {20000/dt_cpu/1e3:.1f} k rows/s; bit-equal: {np.array_equal(want, ids[:20000])}")
    est.kneighbors(Q[:50000])
    t0 = time.perf_counter(); dist, idx = est.kneighbors(Q); dt = time.perf_counter() - t0
    print(f"RFNN kneighbors, raw rows in (fused forest walk + Hamming, n_ref={n_ref}, T={ids.shape[1]}): "
          f"{n_q/dt/1e6:.3f} M queries/s, stats {est.regressor_._get_index().stats()}")


if __name__ == "__main__":
    if "weighted" in sys.argv:
        main(weighted=True)
        main(n_q=200000, T=3500, weighted=True)
    else:
        main()
        forest_main()
