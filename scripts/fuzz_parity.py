"""Randomised parity sweep on the GPU box: random shapes / k / dtypes / duplicates / offsets against
the oracle and against the exhaustive float64 engine (bit-equality between engines).

    python scripts/fuzz_parity.py [n_cases] [seed]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sknnr_oracle as orc
from sknnr_b200 import _lib as L
from sknnr_b200._engine import KNNIndex, HammingIndex


def euclid_case(rng, i):
    d = int(rng.choice([1, 2, 3, 5, 8, 16, 17, 24, 31, 32, 33, 40, 48, 63, 64, 65, 72]))
    n_ref = int(rng.choice([9, 40, 129, 500, 2000, 8191, 8192, 12000, 20001]))
    n_q = int(rng.choice([1, 7, 255, 256, 257, 1000, 3000, 5000]))
    k = int(rng.integers(1, min(n_ref - 1, 24) + 1))
    n_out = int(rng.integers(1, 6))
    R = rng.standard_normal((n_ref, d)) * rng.choice([1.0, 100.0, 1e-3]) + rng.choice([0.0, 50.0])
    dup = rng.random() < 0.4
    if dup and n_ref > 20:
        R[n_ref // 2:n_ref // 2 + n_ref // 10] = R[:n_ref // 10]       # duplicated plots -> exact ties
    y = rng.standard_normal((n_ref, n_out))
    Q = R[rng.integers(0, n_ref, size=n_q)] + rng.standard_normal((n_q, d)) * rng.choice([0.0, 0.1, 1.0]) * R.std()
    mean, scale = R.mean(0), R.std(0, ddof=1)
    scale[scale == 0] = 1.0
    st = orc.FittedState("euclidean", fit_Z=(R - mean) / scale, y=y, center=mean, scale=scale)
    ix = KNNIndex(st.fit_Z, mean, scale, None, y)
    off = int(rng.choice([0, 0, 12345]))
    excl = rng.random() < 0.2 and k + 1 <= n_ref - 1
    Qx = None if excl else Q
    d_o, i_o = orc.kneighbors(st, Qx, k=k, row_offset=0 if excl else off)
    d_g, i_g, p_g = ix.query(Qx, k, exclude_self=excl, weights="distance", with_pred=True,
                             row_offset=0 if excl else off)
    eng = ix.stats()["engine"]
    atol = 1e-7 * float(np.sqrt((st.fit_Z ** 2).sum(1).max())) + 1e-12
    orc.assert_tie_aware_equal(d_g, i_g, d_o, i_o, rtol=1e-5, atol=atol)
    L.set_option("engine", L.ENGINE_EXACT)
    try:
        d_e, i_e, p_e = ix.query(Qx, k, exclude_self=excl, weights="distance", with_pred=True,
                                 row_offset=0 if excl else off)
    finally:
        L.set_option("engine", L.ENGINE_AUTO)
    assert np.array_equal(i_e, i_g) and np.array_equal(d_e, d_g) and np.array_equal(p_e, p_g, equal_nan=True)
    ix.close()
    return f"euclid d={d} n_ref={n_ref} n_q={n_q} k={k} dup={dup} excl={excl} engine={eng} fb={ix.n_out}"


def hamming_case(rng, i):
    T = int(rng.choice([1, 2, 31, 64, 65, 130, 500, 1001]))
    n_ref = int(rng.choice([9, 64, 65, 500, 3000]))
    n_q = int(rng.choice([1, 100, 383, 384, 385, 1500]))
    k = int(rng.integers(1, min(n_ref - 1, 24) + 1))
    n_codes = int(rng.choice([2, 5, 60, 31743]))
    R = rng.integers(0, n_codes, size=(n_ref, T))
    Q = R[rng.integers(0, n_ref, size=n_q)].copy()
    flip = rng.random(Q.shape) < rng.choice([0.0, 0.3, 0.8])
    Q[flip] = rng.integers(0, n_codes, size=int(flip.sum()))
    kind = rng.choice(["eq", "decay", "rand", "zeros"])
    w = {"eq": np.full(T, 1.0 / T), "decay": 0.95 ** np.arange(T), "rand": rng.random(T) + 1e-3,
         "zeros": np.where(rng.random(T) < 0.3, 0.0, rng.random(T))}[kind]
    if w.sum() == 0:
        w[0] = 1.0
    w = w / w.sum()
    st = orc.FittedState("hamming", fit_Z=R, y=np.zeros((n_ref, 1)), hamming_w=w)
    ix = HammingIndex(R.astype(np.uint16), w)
    excl = rng.random() < 0.2 and k + 1 <= n_ref - 1
    Qx = None if excl else Q
    d_o, i_o = orc.kneighbors(st, Qx, k=k)
    d_g, i_g, _ = ix.query(None if excl else Q.astype(np.uint16), k, exclude_self=excl)
    assert np.array_equal(i_g, i_o), "hamming indices"
    assert np.array_equal(d_g, d_o), "hamming distances"
    ix.close()
    return f"hamming T={T} n_ref={n_ref} n_q={n_q} k={k} codes={n_codes} w={kind} excl={excl}"


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    t0 = time.time()
    for i in range(n):
        fn = euclid_case if i % 3 else hamming_case
        try:
            print(i, fn(rng, i), flush=True)
        except Exception as e:  # noqa: BLE001
            print("FAILED case", i, fn.__name__, repr(e)[:300], flush=True)
            raise
    print(f"{n} cases ok in {time.time() - t0:.0f} s")
