"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    PYTHONPATH=/root/reference/src python oracle/make_golden.py

TEST INFRASTRUCTURE ONLY.  The reference is a Python package and cannot travel to the
GPU box, so this script imports it here, fits every estimator on the hot path's scope
list exactly as the reference's own regression tests do
(ref:tests/conftest.py:45-59, ref:tests/test_regressions.py:57-122), and stores

* the inputs (bundled Moscow / SWO data the reference ships),
* the fitted state as flat arrays (centre, scale, projector, transformed reference
  plots, targets, Hamming weights, node-ID matrices),
* the live reference outputs (``live_*``, algorithm="brute"), and
* the reference's own golden vectors for the same case (``refgold_*``, copied from
  ref:tests/test_regressions/*.npz),

so the parity tests can run anywhere.  Library versions are recorded in each file.
"""

from __future__ import annotations

import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))

import scipy  # noqa: E402
import sklearn  # noqa: E402
from sklearn.model_selection import train_test_split  # noqa: E402

import sknnr  # noqa: E402
from sknnr import (  # noqa: E402
    EuclideanKNNRegressor,
    GBNNRegressor,
    GNNRegressor,
    MahalanobisKNNRegressor,
    MSNRegressor,
    RawKNNRegressor,
    RFNNRegressor,
)
from sknnr.datasets import load_moscow_stjoes, load_swo_ecoplot  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
GOLD = os.path.join(REF, "tests", "test_regressions")

VERSIONS = np.array(
    [f"sknnr={sknnr.__version__}", f"sklearn={sklearn.__version__}",
     f"scipy={scipy.__version__}", f"numpy={np.__version__}"]
)


def yaimpute_weights(d):
    # the callable the reference's regression tests use (ref:tests/test_regressions.py:31-39)
    return 1.0 / (1.0 + d)


def affine_state(est):
    """Flatten a fitted float estimator into (center, scale, proj, fit_Z, y)."""
    if isinstance(est, RawKNNRegressor):
        return dict(fit_Z=est._fit_X, y=est._y)
    t = est.transformer_
    name = type(t).__name__
    st = {}
    if name == "StandardScalerWithDOF":
        st.update(center=t.mean_, scale=t.scale_)
    elif name == "MahalanobisTransformer":
        st.update(center=t.scaler_.mean_, scale=t.scaler_.scale_, proj=t.transform_)
    elif name == "CCorATransformer":
        st.update(center=t.scaler_.mean_, scale=t.scaler_.scale_, proj=t.projector_)
    elif name == "CCATransformer":
        st.update(center=t.env_center_, proj=t.projector_)
    else:
        raise TypeError(name)
    st.update(fit_Z=est.regressor_._fit_X, y=est.regressor_._y)
    return st


def load_gold(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        return {}
    with np.load(p) as f:
        return {k: f[k] for k in f.files}


def moscow_split(as_frame=True):
    X, y = load_moscow_stjoes(return_X_y=True, as_frame=as_frame)
    return train_test_split(X, y, train_size=0.8, shuffle=False)


def float_cases():
    Xtr, Xte, ytr, yte = moscow_split()
    np.savez_compressed(
        os.path.join(OUT, "moscow_split.npz"),
        X_train=Xtr.to_numpy(), X_test=Xte.to_numpy(), y_train=ytr.to_numpy(),
        y_test=yte.to_numpy(), index_train=np.asarray(Xtr.index),
        index_test=np.asarray(Xte.index), versions=VERSIONS,
    )
    ests = {
        "raw": RawKNNRegressor, "euclidean": EuclideanKNNRegressor,
        "mahalanobis": MahalanobisKNNRegressor, "gnn": GNNRegressor, "msn": MSNRegressor,
    }
    for name, cls in ests.items():
        for comp_name, n_comp in (("full", None), ("reduced", 3)):
            if n_comp is not None and name not in ("gnn", "msn"):
                continue
            out = {"versions": VERSIONS}
            kw = {"n_neighbors": 5}
            if n_comp is not None:
                kw["n_components"] = n_comp
            est = cls(algorithm="brute", **kw).fit(Xtr, ytr)
            out.update({f"state_{k}": np.asarray(v) for k, v in affine_state(est).items()})
            # live reference outputs (brute)
            d, i = est.kneighbors()
            out["live_ref_dist"], out["live_ref_nn"] = d, i
            out["live_ref_ids"] = est.kneighbors(return_dataframe_index=True)[1]
            d, i = est.kneighbors(Xte)
            out["live_tgt_dist"], out["live_tgt_nn"] = d, i
            out["live_tgt_ids"] = est.kneighbors(Xte, return_dataframe_index=True)[1]
            d, i = est.kneighbors(Xte, use_deterministic_ordering=False)
            out["live_tgt_dist_raw"], out["live_tgt_nn_raw"] = d, i
            out["live_ref_pred_unweighted"] = est.independent_prediction_
            out["live_ref_score_unweighted"] = np.float64(est.independent_score_)
            out["live_tgt_pred_unweighted"] = est.predict(Xte)
            out["live_tgt_score_unweighted"] = np.float64(est.score(Xte, yte))
            for wname, w in (("weighted", yaimpute_weights), ("distance", "distance")):
                ew = cls(algorithm="brute", weights=w, **kw).fit(Xtr, ytr)
                out[f"live_ref_pred_{wname}"] = ew.independent_prediction_
                out[f"live_ref_score_{wname}"] = np.float64(ew.independent_score_)
                out[f"live_tgt_pred_{wname}"] = ew.predict(Xte)
            # the reference's own goldens for this case
            for rt in ("reference", "target"):
                short = "ref" if rt == "reference" else "tgt"
                for ids in ("index", "ids"):
                    g = load_gold(f"test_kneighbors_{rt}_{comp_name}_{name}_k5_{ids}_.npz")
                    for k, v in g.items():
                        out[f"refgold_{short}_{ids}_{k}"] = v
                for wname in ("weighted", "unweighted"):
                    g = load_gold(f"test_predict_{rt}_{wname}_{comp_name}_{name}_k5_.npz")
                    for k, v in g.items():
                        out[f"refgold_{short}_{wname}_{k}"] = v
            np.savez_compressed(os.path.join(OUT, f"moscow_{name}_{comp_name}.npz"), **out)
            print("wrote", name, comp_name, sorted(k for k in out if k.startswith("refgold"))[:3])


def config_cases():
    # C1: MSNRegressor(n_neighbors=5) on SWO Ecoplot
    X, y = load_swo_ecoplot(return_X_y=True, as_frame=True)
    est = MSNRegressor(n_neighbors=5, algorithm="brute").fit(X, y)
    out = {"versions": VERSIONS, "X": X.to_numpy(), "y_targets": y.to_numpy(),
           "index": np.asarray(X.index)}
    out.update({f"state_{k}": np.asarray(v) for k, v in affine_state(est).items()})
    out["live_ref_dist"], out["live_ref_nn"] = est.kneighbors()
    out["live_self_dist"], out["live_self_nn"] = est.kneighbors(X)
    out["live_ref_pred"] = est.independent_prediction_
    out["live_ref_score"] = np.float64(est.independent_score_)
    out["live_self_pred"] = est.predict(X)
    ed = MSNRegressor(n_neighbors=5, algorithm="brute", weights="distance").fit(X, y)
    out["live_self_pred_distance"] = ed.predict(X)
    out["live_ref_pred_distance"] = ed.independent_prediction_
    np.savez_compressed(os.path.join(OUT, "c1_swo_msn_k5.npz"), **out)
    print("wrote c1", est.regressor_._fit_X.shape)

    # C2: GNNRegressor on the full Moscow data with independent_score_
    X, y = load_moscow_stjoes(return_X_y=True, as_frame=True)
    est = GNNRegressor(n_neighbors=5, algorithm="brute").fit(X, y)
    out = {"versions": VERSIONS, "X": X.to_numpy(), "y_targets": y.to_numpy(),
           "index": np.asarray(X.index)}
    out.update({f"state_{k}": np.asarray(v) for k, v in affine_state(est).items()})
    out["live_ref_dist"], out["live_ref_nn"] = est.kneighbors()
    out["live_ref_pred"] = est.independent_prediction_
    out["live_ref_score"] = np.float64(est.independent_score_)
    np.savez_compressed(os.path.join(OUT, "c2_moscow_gnn_k5.npz"), **out)
    print("wrote c2", est.regressor_._fit_X.shape, est.independent_score_)


def rfnn_case():
    Xtr, Xte, ytr, yte = moscow_split()
    est = RFNNRegressor(n_neighbors=5, random_state=42).fit(Xtr, ytr)
    ids_tr = est.transformer_.transform(Xtr)
    ids_te = est.transformer_.transform(Xte)
    assert ids_tr.max() < 2**15
    out = {"versions": VERSIONS, "ids_train": ids_tr.astype(np.int16),
           "ids_test": ids_te.astype(np.int16), "hamming_w": est.hamming_weights_,
           "y": est.regressor_._y}
    out["live_ref_dist"], out["live_ref_nn"] = est.kneighbors()
    out["live_tgt_dist"], out["live_tgt_nn"] = est.kneighbors(Xte)
    out["live_ref_pred"] = est.independent_prediction_
    out["live_ref_score"] = np.float64(est.independent_score_)
    out["live_tgt_pred"] = est.predict(Xte)
    # non-uniform forest weights (user-supplied), same forests
    fw = np.linspace(1.0, 3.0, ytr.shape[1])
    est2 = RFNNRegressor(n_neighbors=5, random_state=42, forest_weights=fw).fit(Xtr, ytr)
    assert np.array_equal(est2.transformer_.transform(Xte), ids_te)
    out["hamming_w_nonuniform"] = est2.hamming_weights_
    out["live_tgt_dist_nonuniform"], out["live_tgt_nn_nonuniform"] = est2.kneighbors(Xte)
    np.savez_compressed(os.path.join(OUT, "moscow_rfnn.npz"), **out)
    print("wrote rfnn", ids_tr.shape, ids_tr.max())


def mixed_y_fit(ytr):
    # regression + 3-class classification targets of the reference's mixed-forest test
    # (ref:tests/test_regressions.py:158-176)
    cols = [c for c in ytr.columns if c.endswith("_BA") and c != "Total_BA"]
    mx = ytr[cols].idxmax(axis=1)
    mx = mx.where(mx.isin(["ABGR_BA", "TSHE_BA"]), other="OTHER")
    return ytr[["Total_BA"]].assign(MAX_SPECIES=mx)


def flat_trees(transformer):
    """scikit-learn tree_ arrays of every tree in transform's column order, concatenated."""
    trees = []
    for est in transformer.estimators_:
        e = est.estimators_
        for c in range(e.shape[1]):
            for s_ in range(e.shape[0]):
                trees.append(e[s_, c].tree_)
    offs = np.zeros(len(trees) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([t.node_count for t in trees])
    cat = lambda n, dt: np.concatenate([np.asarray(getattr(t, n)) for t in trees]).astype(dt)  # noqa: E731
    return {"tree_offsets": offs, "tree_left": cat("children_left", np.int32),
            "tree_right": cat("children_right", np.int32), "tree_feature": cat("feature", np.int32),
            "tree_threshold": cat("threshold", np.float64)}


def gbnn_case():
    """GBNNRegressor (scope row f3): train-improvement tree weights make the Hamming weights
    unequal; the mixed case adds a 3-class GradientBoostingClassifier (3 trees per stage)."""
    Xtr, Xte, ytr, yte = moscow_split()
    out = {"versions": VERSIONS}
    yf = mixed_y_fit(ytr)
    out["mixed_yfit_total_ba"] = yf["Total_BA"].to_numpy()
    out["mixed_yfit_max_species"] = yf["MAX_SPECIES"].to_numpy().astype("U8")
    for tag, y_fit in (("", None), ("mixed_", yf)):
        est = GBNNRegressor(n_neighbors=5, random_state=42).fit(Xtr, ytr, y_fit=y_fit)
        ids_tr = est.transformer_.transform(Xtr)
        ids_te = est.transformer_.transform(Xte)
        assert ids_tr.max() < 2**15 and ids_te.max() < 2**15
        out[tag + "ids_train"] = ids_tr.astype(np.int16)
        out[tag + "ids_test"] = ids_te.astype(np.int16)
        out[tag + "hamming_w"] = est.hamming_weights_
        out[tag + "tree_weights"] = np.hstack(est.transformer_.tree_weights_)
        out[tag + "n_trees_per_iteration"] = np.asarray(est.transformer_.n_trees_per_iteration_)
        out[tag + "y"] = est.regressor_._y
        for k, v in flat_trees(est.transformer_).items():
            out[tag + k] = v
        out[tag + "live_ref_dist"], out[tag + "live_ref_nn"] = est.kneighbors()
        out[tag + "live_tgt_dist"], out[tag + "live_tgt_nn"] = est.kneighbors(Xte)
        out[tag + "live_ref_pred"] = est.independent_prediction_
        out[tag + "live_ref_score"] = np.float64(est.independent_score_)
        out[tag + "live_tgt_pred"] = est.predict(Xte)
    # the reference's own golden vectors for these cases
    for rt, short in (("reference", "ref"), ("target", "tgt")):
        for k, v in load_gold(f"test_kneighbors_{rt}_full_gbnn_k5_index_.npz").items():
            out[f"refgold_{short}_index_{k}"] = v
        for k, v in load_gold(f"test_predict_{rt}_unweighted_full_gbnn_k5_.npz").items():
            out[f"refgold_{short}_unweighted_{k}"] = v
        for k, v in load_gold(f"test_estimators_with_mixed_type_forests_{rt}_gbnn_.npz").items():
            out[f"refgold_mixed_{short}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "moscow_gbnn.npz"), **out)
    print("wrote gbnn", out["ids_train"].shape, out["mixed_ids_train"].shape,
          sorted(k for k in out if k.startswith("refgold")))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cases = {"float": float_cases, "config": config_cases, "rfnn": rfnn_case, "gbnn": gbnn_case}
    for name in (sys.argv[1:] or list(cases)):
        cases[name]()
