"""CPU oracle for sknnr's query-time hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  Nothing under ``sknnr_b200/`` imports it, and the product path
raises when the CUDA library is missing instead of falling back to this code.

It restates, in plain NumPy (plus the small C file ``hamming_oracle.c`` for the
sequential-sum weighted Hamming distance), the algorithm the reference runs for

    projection -> brute pairwise distance -> per-query top-k -> weighted average

``ref:`` below is ``/root/reference/`` (lemma-osu/sknnr v0.1.0a3).  The arithmetic of
that path lives in third-party code that is NOT under ``/root/reference``:
scikit-learn (declared ``>=1.6.0``, unpinned; 1.9.0 in this image), SciPy (unpinned;
1.18.1 here) and NumPy (2.3.5 here).  ``$SP`` = site-packages of this image.  Those
algorithms are restated from their published sources and pinned (see
``tests/test_oracle.py``) against

* the reference's own golden vectors ``ref:tests/test_regressions/*.npz`` (float
  estimators reproduce exactly with scikit-learn 1.9.0), and
* outputs of the unmodified reference run in the build container by
  ``oracle/make_golden.py`` (committed under ``tests/golden/``).

Parity status: PINNED for the Euclidean-space estimators (raw / euclidean /
mahalanobis / msn / gnn) and for the Hamming path given shared node-ID matrices; the
reference's random-forest goldens do not reproduce under scikit-learn 1.9.0 (forest
training differs), so RFNN is pinned on live-reference outputs only.
"""

from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------
# (a1-a4) projection into the estimator's feature space
# ----------------------------------------------------------------------------------
def affine_project(X, center=None, scale=None, proj=None):
    """``Z = ((X - center) / scale) @ proj`` in float64.

    One formula covers the four float transformers:

    * StandardScalerWithDOF  - ``(X - mean_) / scale_``
      ($SP/sklearn/preprocessing/_data.py:1131-1134; scale_ set at
      ref:src/sknnr/transformers/_base.py:66)
    * MahalanobisTransformer - ``scaler_.transform(X) @ transform_``
      (ref:src/sknnr/transformers/_mahalanobis_transformer.py:55)
    * CCorATransformer       - ``scaler_.transform(X) @ projector_``
      (ref:src/sknnr/transformers/_ccora_transformer.py:70)
    * CCATransformer         - ``(X - env_center_) @ projector_`` (no scaling)
      (ref:src/sknnr/transformers/_cca_transformer.py:87)
    """
    Z = np.array(X, dtype=np.float64, copy=True)
    if center is not None:
        Z -= np.asarray(center, dtype=np.float64)
    if scale is not None:
        Z /= np.asarray(scale, dtype=np.float64)
    if proj is not None:
        Z = Z @ np.asarray(proj, dtype=np.float64)
    return Z


# ----------------------------------------------------------------------------------
# (a7) brute Euclidean k-nearest neighbours == sklearn EuclideanArgKmin64
# ----------------------------------------------------------------------------------
def brute_kneighbors_euclidean(Z, fit_Z, k, chunk=2048):
    """Restatement of scikit-learn's ``EuclideanArgKmin64``.

    $SP/sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:311-510 and
    _middle_term_computer.pyx.tp:401-441: squared distances are formed in float64 from
    the expansion ``|x|^2 - 2 x.y + |y|^2`` (:494-502), clamped at 0, pushed into a
    per-query max-heap that rejects ``val >= heap top`` ($SP/sklearn/utils/_heap.pyx:46)
    while references are scanned in ascending index order - so among equal distances at
    the k-th boundary the LOWEST index is kept - then each heap is sorted ascending and
    ``sqrt`` is taken of the expansion value (:285-295).  Called from
    ``KNeighborsMixin.kneighbors`` ($SP/sklearn/neighbors/_base.py:855-870), which sknnr
    reaches at ref:src/sknnr/_base.py:162-164.

    A stable argsort on the clamped expansion reproduces "ascending, lowest index among
    equals".  (The dgemm summation order is BLAS-internal; values agree to ~1 ulp.)
    """
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    Y = np.ascontiguousarray(fit_Z, dtype=np.float64)
    n_q = Z.shape[0]
    yn = np.einsum("ij,ij->i", Y, Y)
    dist = np.empty((n_q, k), dtype=np.float64)
    idx = np.empty((n_q, k), dtype=np.int64)
    for s in range(0, n_q, chunk):
        x = Z[s : s + chunk]
        xn = np.einsum("ij,ij->i", x, x)
        d2 = xn[:, None] - 2.0 * (x @ Y.T) + yn[None, :]
        np.maximum(d2, 0.0, out=d2)
        if k < d2.shape[1]:
            # k smallest with lowest-index tie-break at the boundary: take everything
            # strictly below the k-th value plus the lowest-index members equal to it.
            part = np.partition(d2, k - 1, axis=1)[:, k - 1]
            order = np.empty((d2.shape[0], k), dtype=np.int64)
            for r in range(d2.shape[0]):
                row = d2[r]
                cand = np.flatnonzero(row <= part[r])
                o = cand[np.argsort(row[cand], kind="stable")][:k]
                order[r] = o
        else:
            order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        rows = np.arange(d2.shape[0])[:, None]
        dist[s : s + chunk] = np.sqrt(d2[rows, order])
        idx[s : s + chunk] = order
    return dist, idx


# ----------------------------------------------------------------------------------
# (a8) weighted Hamming distance == scipy cdist_hamming
# ----------------------------------------------------------------------------------
_hamming_lib = None


def _load_hamming_lib():
    global _hamming_lib
    if _hamming_lib is None:
        path = os.path.join(_HERE, "_build", "libhamming_oracle.so")
        if not os.path.exists(path):
            build_c()
        lib = ctypes.CDLL(path)
        lib.hamming_cdist_w.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
        ]
        lib.hamming_cdist_w.restype = None
        _hamming_lib = lib
    return _hamming_lib


def build_c():
    """Compile ``hamming_oracle.c`` into ``oracle/_build/`` (called by ``build()``)."""
    import subprocess

    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libhamming_oracle.so")
    src = os.path.join(_HERE, "hamming_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(
        ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-ffp-contract=off", "-o", out, src]
    )
    return out


def hamming_lut(w):
    """Distances reachable with EQUAL weights: ``lut[m]`` = distance at m mismatches.

    SciPy 1.18.1's weighted Hamming ($SP/scipy/spatial/distance.py:1718-1723 ->
    ``_distance_pybind.cdist_hamming``) is, bit for bit, ``(sum of w_t over mismatching
    trees, accumulated strictly left to right in float64) / (sum of all w_t, same
    order)``.  With equal w_t the numerator depends only on the mismatch COUNT.
    """
    w = np.asarray(w, dtype=np.float64)
    den = 0.0
    for v in w:
        den = den + float(v)
    lut = np.empty(w.shape[0] + 1, dtype=np.float64)
    acc = 0.0
    lut[0] = 0.0
    for m in range(1, w.shape[0] + 1):
        acc = acc + float(w[0])
        lut[m] = acc / den
    return lut


def hamming_cdist(Q_ids, R_ids, w, use_c=True):
    """Full weighted-Hamming distance matrix, float64 ``[n_q, n_ref]``.

    Follows the call made by sklearn's brute fallback for ``metric="hamming"``
    ($SP/sklearn/neighbors/_base.py:879-908 -> pairwise_distances_chunked ->
    scipy cdist) with ``w = hamming_weights_`` (ref:src/sknnr/_weighted_trees.py:65-98,
    passed at :139-140).
    """
    Q = np.ascontiguousarray(Q_ids, dtype=np.int64)
    R = np.ascontiguousarray(R_ids, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    n_q, T = Q.shape
    n_r = R.shape[0]
    out = np.empty((n_q, n_r), dtype=np.float64)
    if use_c:
        lib = _load_hamming_lib()
        lib.hamming_cdist_w(
            Q.ctypes.data, R.ctypes.data, w.ctypes.data, n_q, n_r, T, out.ctypes.data
        )
        return out
    den = 0.0
    for v in w:
        den = den + float(v)
    out[:] = 0.0
    for t in range(T):  # strictly left-to-right accumulation, vectorised over pairs
        out += np.where(Q[:, t, None] != R[None, :, t], w[t], 0.0)
    out /= den
    return out


def brute_kneighbors_hamming(Q_ids, R_ids, w, k):
    """CANONICAL Hamming oracle: full cdist, then k smallest by (distance, index).

    The reference reduces each distance slab with ``np.argpartition`` + ``argsort``
    ($SP/sklearn/neighbors/_base.py:743-746), whose choice among equal distances at the
    k-th boundary is introselect-arbitrary.  The canonical form keeps the LOWEST indices
    among boundary ties - what a deterministic GPU top-k must equal bit-exactly - and is
    compared with the raw reference through ``assert_tie_aware_equal``.
    """
    D = hamming_cdist(Q_ids, R_ids, w)
    order = np.argsort(D, axis=1, kind="stable")[:, :k]
    rows = np.arange(D.shape[0])[:, None]
    return D[rows, order], order.astype(np.int64)


# ----------------------------------------------------------------------------------
# (a9) X=None: drop the query itself from k+1 neighbours
# ----------------------------------------------------------------------------------
def exclude_self(dist, idx, row_offset=0):
    """$SP/sklearn/neighbors/_base.py:929-958: delete the column whose index equals the
    query's row; when no column matches (>= k+1 exact duplicates) delete column 0."""
    n_q, k1 = idx.shape
    rows = np.arange(row_offset, row_offset + n_q)[:, None]
    mask = idx != rows
    dup = np.all(mask, axis=1)
    mask[:, 0][dup] = False
    return dist[mask].reshape(n_q, k1 - 1), idx[mask].reshape(n_q, k1 - 1)


# ----------------------------------------------------------------------------------
# (a6) sknnr's deterministic neighbour ordering
# ----------------------------------------------------------------------------------
def deterministic_order(dist, idx, decimals=10, row_offset=0):
    """ref:src/sknnr/_base.py:166-175: re-sort the k columns by
    ``(round(dist / max(rowmax, 1), decimals), |idx - query_row|, idx)``.

    ``query_row`` is the row position within the batch passed to the call (:171);
    ``row_offset`` lets a sharded caller reproduce the single-call result.
    """
    row_scale = np.maximum(dist.max(axis=1, keepdims=True), 1.0)
    rounded = np.round(dist / row_scale, decimals=decimals)
    rows = np.arange(row_offset, row_offset + len(idx))[:, None]
    diff = np.abs(idx - rows)
    order = np.lexsort((idx, diff, rounded), axis=1)
    return np.take_along_axis(dist, order, axis=1), np.take_along_axis(idx, order, axis=1)


# ----------------------------------------------------------------------------------
# (a10) prediction weights and the multi-output average
# ----------------------------------------------------------------------------------
def get_weights(dist, weights):
    """$SP/sklearn/neighbors/_base.py:74-117 (``_get_weights``)."""
    if weights in (None, "uniform"):
        return None
    if isinstance(weights, str) and weights == "distance":
        with np.errstate(divide="ignore"):
            w = 1.0 / dist
        inf_mask = np.isinf(w)
        inf_row = np.any(inf_mask, axis=1)
        w[inf_row] = inf_mask[inf_row]
        return w
    if callable(weights):
        return weights(dist)
    raise ValueError(weights)


def weighted_average(y, idx, w=None):
    """$SP/sklearn/neighbors/_regression.py:254-267."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y.reshape(-1, 1)
    if w is None:
        return np.mean(y[idx], axis=1)
    pred = np.empty((idx.shape[0], y.shape[1]), dtype=np.float64)
    denom = np.sum(w, axis=1)
    for j in range(y.shape[1]):
        pred[:, j] = np.sum(y[idx, j] * w, axis=1) / denom
    return pred


def r2_score_uniform(y_true, y_pred):
    """``RegressorMixin.score`` -> ``r2_score(multioutput='uniform_average')``
    ($SP/sklearn/base.py:672-716), reached from ref:src/sknnr/_base.py:40,350-352."""
    y_true = np.asarray(y_true, dtype=np.float64)
    y_pred = np.asarray(y_pred, dtype=np.float64)
    if y_true.ndim == 1:
        y_true = y_true[:, None]
        y_pred = y_pred.reshape(-1, 1)
    num = ((y_true - y_pred) ** 2).sum(axis=0)
    den = ((y_true - y_true.mean(axis=0)) ** 2).sum(axis=0)
    nz_den = den != 0
    nz_num = num != 0
    out = np.ones(y_true.shape[1])
    valid = nz_den & nz_num
    out[valid] = 1 - num[valid] / den[valid]
    out[nz_num & ~nz_den] = 0.0
    return float(out.mean())


# ----------------------------------------------------------------------------------
# fitted state + end-to-end path
# ----------------------------------------------------------------------------------
@dataclass
class FittedState:
    """Flat arrays that fully determine the query-time path of one fitted estimator."""

    kind: str  # "euclidean" | "hamming"
    fit_Z: np.ndarray  # [n_ref, d'] float64, or [n_ref, T] int64 node IDs
    y: np.ndarray  # [n_ref, n_out]
    center: np.ndarray | None = None
    scale: np.ndarray | None = None
    proj: np.ndarray | None = None
    hamming_w: np.ndarray | None = None
    extra: dict = field(default_factory=dict)


def transform(state: FittedState, X):
    if state.kind == "hamming":
        return np.asarray(X, dtype=np.int64)  # node IDs are supplied by the caller
    return affine_project(X, state.center, state.scale, state.proj)


def kneighbors(
    state: FittedState,
    X=None,
    k=5,
    deterministic=True,
    decimals=10,
    row_offset=0,
    transformed=False,
):
    """ref:src/sknnr/_base.py:285-344 -> :111-182 (whole kneighbors call)."""
    query_is_train = X is None
    if query_is_train:
        Z = state.fit_Z
        kk = k + 1
    else:
        Z = X if transformed else transform(state, X)
        kk = k
    if state.kind == "hamming":
        dist, idx = brute_kneighbors_hamming(Z, state.fit_Z, state.hamming_w, kk)
    else:
        dist, idx = brute_kneighbors_euclidean(Z, state.fit_Z, kk)
    if query_is_train:
        dist, idx = exclude_self(dist, idx, row_offset)
    if deterministic:
        dist, idx = deterministic_order(dist, idx, decimals, row_offset)
    return dist, idx


def predict(state: FittedState, X=None, k=5, weights="uniform", **kw):
    """ref:src/sknnr/_base.py:346-348 -> $SP/sklearn/neighbors/_regression.py:229-273."""
    dist, idx = kneighbors(state, X, k, **kw)
    return weighted_average(state.y, idx, get_weights(dist, weights))


# ----------------------------------------------------------------------------------
# raster caller: what a map-producing caller does around the path
# ----------------------------------------------------------------------------------
def raster_query(state: FittedState, image, k=5, nodata=None, weights=None, deterministic=True,
                 decimals=10, fill_dist=np.nan, fill_idx=-1, fill_pred=np.nan):
    """The host-side loop around ref:src/sknnr/_base.py:285-352 for a band-major image
    ``[bands, H, W]``: flatten to ``[H*W, bands]`` rows, keep the pixels whose bands are all finite
    and different from ``nodata``, query them in pixel order as ONE kneighbors / predict call, and
    scatter the results back into ``[k, H, W]`` / ``[n_out, H, W]`` layers (fill where masked)."""
    image = np.asarray(image)
    nb, hw = image.shape[0], image.shape[1:]
    X = image.reshape(nb, -1).T
    valid = np.isfinite(X.astype(np.float64)).all(axis=1)
    if nodata is not None:
        valid &= ~(X == nodata).any(axis=1)
    n_pix = X.shape[0]
    dist = np.full((k, n_pix), fill_dist, dtype=np.float64)
    idx = np.full((k, n_pix), fill_idx, dtype=np.int64)
    pred = None
    if weights is not None:
        pred = np.full((state.y.shape[1], n_pix), fill_pred, dtype=np.float64)
    if valid.any():
        d, i = kneighbors(state, np.asarray(X[valid], dtype=np.float64), k, deterministic, decimals)
        dist[:, valid] = d.T
        idx[:, valid] = i.T
        if weights is not None:
            pred[:, valid] = weighted_average(state.y, i, get_weights(d, weights)).T
    out = (dist.reshape(k, *hw), idx.reshape(k, *hw))
    return out + ((pred.reshape(-1, *hw),) if weights is not None else ())


# ----------------------------------------------------------------------------------
# comparators
# ----------------------------------------------------------------------------------
def assert_tie_aware_equal(dist, idx, ref_dist, ref_idx, rtol=1e-5, atol=1e-9, gap_rtol=2e-6):
    """Per row: distances must agree; an index may differ from the reference only inside
    a group of (near-)equal distances (relative gap < ``gap_rtol``)."""
    dist = np.asarray(dist)
    ref_dist = np.asarray(ref_dist)
    np.testing.assert_allclose(dist, ref_dist, rtol=rtol, atol=atol)
    bad = np.flatnonzero((idx != ref_idx).any(axis=1))
    for r in bad:
        a, b = set(idx[r].tolist()), set(ref_idx[r].tolist())
        scale = max(float(ref_dist[r].max()), 1e-300)
        for c in np.flatnonzero(idx[r] != ref_idx[r]):
            d = ref_dist[r, c]
            near = np.abs(ref_dist[r] - d) <= gap_rtol * scale + atol
            same_group = int(near.sum()) > 1 or c == idx.shape[1] - 1
            if not same_group:
                raise AssertionError(
                    f"row {r}: index {idx[r, c]} != {ref_idx[r, c]} at col {c} with no "
                    f"tie (dist {dist[r]} vs {ref_dist[r]}; only-in-ours {a - b})"
                )
    return len(bad)


# ----------------------------------------------------------------------------------
# CPU baseline legs for bench.py (the arithmetic the reference itself executes)
# ----------------------------------------------------------------------------------
def sklearn_reference_kneighbors(fit_Z, y, Z, k, metric="euclidean", w=None, n_jobs=None,
                                 deterministic=True):
    """Time-able stand-in for ``RawKNNRegressor.kneighbors``: scikit-learn's own
    ``KNeighborsRegressor(algorithm='brute')`` (the third-party kernel sknnr delegates
    to at ref:src/sknnr/_base.py:162-164) followed by the sknnr re-ordering glue
    restated above.  Returns (dist, idx, fitted sklearn regressor)."""
    from sklearn.neighbors import KNeighborsRegressor

    kw = dict(n_neighbors=k, algorithm="brute", n_jobs=n_jobs)
    if metric == "hamming":
        kw.update(metric="hamming", metric_params={"w": w})
    reg = KNeighborsRegressor(**kw).fit(fit_Z, y)
    dist, idx = reg.kneighbors(Z)
    if deterministic:
        dist, idx = deterministic_order(dist, idx)
    return dist, idx, reg
