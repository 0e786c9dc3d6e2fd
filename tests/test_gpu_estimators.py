"""GPU tests of the estimator surface (drop-in behaviour) against the reference's goldens."""

import pickle

import numpy as np
import pytest

from oracle import sknnr_oracle as orc
from tests.conftest import load_golden, yaimpute_weights

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _split(as_frame=False):
    sp = load_golden("moscow_split.npz")
    if not as_frame:
        return sp["X_train"], sp["X_test"], sp["y_train"], sp["y_test"], sp
    import pandas as pd

    cols = [f"f{i}" for i in range(sp["X_train"].shape[1])]
    Xtr = pd.DataFrame(sp["X_train"], columns=cols, index=sp["index_train"])
    Xte = pd.DataFrame(sp["X_test"], columns=cols, index=sp["index_test"])
    return Xtr, Xte, sp["y_train"], sp["y_test"], sp


def _estimators():
    import sknnr_b200 as S

    return {"raw": S.RawKNNRegressor, "euclidean": S.EuclideanKNNRegressor,
            "mahalanobis": S.MahalanobisKNNRegressor, "gnn": S.GNNRegressor, "msn": S.MSNRegressor}


CASES = [("raw", None), ("euclidean", None), ("mahalanobis", None), ("gnn", None), ("gnn", 3),
         ("msn", None), ("msn", 3)]


@pytest.mark.parametrize(("name", "n_comp"), CASES)
def test_regression_goldens_through_estimators(name, n_comp):
    """ref:tests/test_regressions.py:57-122 replayed on the drop-in classes."""
    g = load_golden(f"moscow_{name}_{'reduced' if n_comp else 'full'}.npz")
    Xtr, Xte, ytr, yte, sp = _split(as_frame=True)
    kw = {"n_neighbors": 5}
    if n_comp:
        kw["n_components"] = n_comp
    atol = 1e-7 * float(np.sqrt((g["state_fit_Z"] ** 2).sum(1).max())) + 1e-12
    est = _estimators()[name](**kw).fit(Xtr, ytr)
    # kneighbors: reference (X=None) and target, array index and dataframe ids
    for ids in (False, True):
        key = "ids" if ids else "index"
        d, i = est.kneighbors(return_dataframe_index=ids)
        np.testing.assert_array_equal(i, g[f"refgold_ref_{key}_nn"])
        np.testing.assert_allclose(d, g[f"refgold_ref_{key}_dist"], rtol=RTOL, atol=atol)
        d, i = est.kneighbors(Xte, return_dataframe_index=ids)
        np.testing.assert_array_equal(i, g[f"refgold_tgt_{key}_nn"])
        np.testing.assert_allclose(d, g[f"refgold_tgt_{key}_dist"], rtol=RTOL, atol=atol)
    assert est.kneighbors(Xte, return_distance=False).shape == (33, 5)
    # predict / independent prediction / score, unweighted and yaImpute-weighted
    np.testing.assert_allclose(est.independent_prediction_, g["refgold_ref_unweighted_pred"], rtol=RTOL, atol=1e-8)
    assert est.independent_score_ == pytest.approx(float(g["refgold_ref_unweighted_score"]), abs=1e-6)
    np.testing.assert_allclose(est.predict(Xte), g["refgold_tgt_unweighted_pred"], rtol=RTOL, atol=1e-8)
    assert est.score(Xte, yte) == pytest.approx(float(g["live_tgt_score_unweighted"]), abs=1e-6)
    ew = _estimators()[name](weights=yaimpute_weights, **kw).fit(Xtr, ytr)
    np.testing.assert_allclose(ew.independent_prediction_, g["refgold_ref_weighted_pred"], rtol=RTOL, atol=1e-8)
    assert ew.independent_score_ == pytest.approx(float(g["refgold_ref_weighted_score"]), abs=1e-6)
    np.testing.assert_allclose(ew.predict(Xte), g["refgold_tgt_weighted_pred"], rtol=RTOL, atol=1e-8)
    ed = _estimators()[name](weights="distance", **kw).fit(Xtr, ytr)
    np.testing.assert_allclose(ed.predict(Xte), g["live_tgt_pred_distance"], rtol=RTOL, atol=1e-8)
    # fitted attributes of the reference surface
    assert est.n_features_in_ == g["state_fit_Z"].shape[1]
    np.testing.assert_array_equal(est.dataframe_index_in_, sp["index_train"])
    if name != "raw":
        np.testing.assert_allclose(est.regressor_._fit_X, g["state_fit_Z"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(est.transformer_.transform(Xte), orc.affine_project(
            sp["X_test"], g.get("state_center"), g.get("state_scale"), g.get("state_proj")), rtol=1e-9, atol=1e-9)


def test_config1_msn_swo_through_estimator():
    import sknnr_b200 as S

    g = load_golden("c1_swo_msn_k5.npz")
    est = S.MSNRegressor(n_neighbors=5).fit(g["X"], g["y_targets"])
    d, i = est.kneighbors()
    orc.assert_tie_aware_equal(d, i, g["live_ref_dist"], g["live_ref_nn"], rtol=RTOL, atol=1e-6)
    assert est.independent_score_ == pytest.approx(float(g["live_ref_score"]), abs=1e-6)
    np.testing.assert_allclose(est.predict(g["X"][:500]), g["live_self_pred"][:500], rtol=RTOL, atol=1e-6)


def test_config2_gnn_moscow_independent_score():
    import sknnr_b200 as S

    g = load_golden("c2_moscow_gnn_k5.npz")
    est = S.GNNRegressor(n_neighbors=5).fit(g["X"], g["y_targets"])
    assert est.independent_score_ == pytest.approx(float(g["live_ref_score"]), abs=1e-6)
    np.testing.assert_allclose(est.independent_prediction_, g["live_ref_pred"], rtol=RTOL, atol=1e-8)
    d, i = est.kneighbors()
    np.testing.assert_array_equal(i, g["live_ref_nn"])


def _same_forests_or_skip(g, ids_te, key, what):
    """The golden node IDs come from forests grown by a recorded scikit-learn / NumPy: with the same
    versions different forests are a FAILURE (the seeded training is deterministic), with other
    versions the end-to-end comparison cannot be made and the test says so loudly."""
    import numpy
    import sklearn

    if np.array_equal(ids_te, g[key].astype(np.int64)):
        return
    have = {"sklearn": sklearn.__version__, "numpy": numpy.__version__}
    rec = dict(v.split("=", 1) for v in g["versions"].tolist())
    same = all(rec.get(k) == v for k, v in have.items())
    assert not same, f"{what}: different trees although the versions equal the golden generator's ({rec})"
    pytest.skip(f"{what}: scikit-learn {have['sklearn']} / NumPy {have['numpy']} grew different trees than the "
                f"golden generator's ({rec.get('sklearn')} / {rec.get('numpy')})")


def test_rfnn_end_to_end_matches_live_reference():
    """Same scikit-learn version and seed grow the same forests, so the whole RFNN path
    (forest apply on the host, Hamming search on the device) must reproduce the live
    reference: bit-equal distances, indices equal up to the reference's arbitrary boundary
    ties (canonical lowest-index order on our side)."""
    import sknnr_b200 as S

    g = load_golden("moscow_rfnn.npz")
    Xtr, Xte, ytr, yte, _ = _split()
    est = S.RFNNRegressor(n_neighbors=5, random_state=42).fit(Xtr, ytr)
    ids_te = est.transformer_.transform(Xte)
    _same_forests_or_skip(g, ids_te, "ids_test", "RFNN end to end")
    np.testing.assert_array_equal(est.hamming_weights_, g["hamming_w"])
    d, i = est.kneighbors(Xte)
    assert np.array_equal(d, g["live_tgt_dist"])
    orc.assert_tie_aware_equal(d, i, g["live_tgt_dist"], g["live_tgt_nn"], rtol=0, atol=0, gap_rtol=0)
    d, i = est.kneighbors()
    assert np.array_equal(d, g["live_ref_dist"])
    st = orc.FittedState("hamming", fit_Z=g["ids_train"].astype(np.int64), y=g["y"], hamming_w=g["hamming_w"])
    d_o, i_o = orc.kneighbors(st, None, k=5)
    np.testing.assert_array_equal(i, i_o)
    assert est.independent_prediction_.shape == g["live_ref_pred"].shape
    # user-supplied forest weights -> unequal tree weights -> exact float64 kernel
    fw = np.linspace(1.0, 3.0, ytr.shape[1])
    est2 = S.RFNNRegressor(n_neighbors=5, random_state=42, forest_weights=fw).fit(Xtr, ytr)
    np.testing.assert_array_equal(est2.hamming_weights_, g["hamming_w_nonuniform"])
    d, i = est2.kneighbors(Xte)
    assert np.array_equal(d, g["live_tgt_dist_nonuniform"])


def test_gbnn_end_to_end_matches_live_reference():
    """GBNNRegressor (scope row f3) end to end - boosted models trained by scikit-learn, forest walk
    and weighted-Hamming search on the device - against the live reference: bit-equal distances,
    indices equal up to boundary ties, same predictions and score."""
    import warnings

    import pandas as pd

    import sknnr_b200 as S

    g = load_golden("moscow_gbnn.npz")
    Xtr, Xte, ytr, yte, _ = _split()
    y_fit_mixed = pd.DataFrame({"Total_BA": g["mixed_yfit_total_ba"], "MAX_SPECIES": g["mixed_yfit_max_species"]})
    for tag, y_fit in (("", None), ("mixed_", y_fit_mixed)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", FutureWarning)
            est = S.GBNNRegressor(n_neighbors=5, random_state=42).fit(Xtr, ytr, y_fit=y_fit)
        ids_te = est.transformer_.transform(Xte)
        _same_forests_or_skip(g, ids_te, tag + "ids_test", "GBNN end to end")
        np.testing.assert_allclose(est.hamming_weights_, g[tag + "hamming_w"], rtol=1e-12)
        if not np.array_equal(est.hamming_weights_, g[tag + "hamming_w"]):
            continue   # weights differ in the last bit: distances cannot be bit-equal
        d, i = est.kneighbors(Xte)
        assert np.array_equal(d, g[tag + "live_tgt_dist"])
        orc.assert_tie_aware_equal(d, i, g[tag + "live_tgt_dist"], g[tag + "live_tgt_nn"], rtol=0, atol=0, gap_rtol=0)
        d, i = est.kneighbors()
        assert np.array_equal(d, g[tag + "live_ref_dist"])
        np.testing.assert_allclose(est.predict(Xte), g[tag + "live_tgt_pred"], rtol=1e-12)
        np.testing.assert_allclose(est.independent_prediction_, g[tag + "live_ref_pred"], rtol=1e-12)
        assert est.independent_score_ == pytest.approx(float(g[tag + "live_ref_score"]), abs=1e-12)


def test_estimator_hygiene_pickle_lists_gridsearch_and_errors():
    import sknnr_b200 as S
    from sklearn.model_selection import GridSearchCV

    Xtr, Xte, ytr, yte, _ = _split()
    est = S.EuclideanKNNRegressor(n_neighbors=3).fit(Xtr.tolist(), ytr.tolist())
    assert not hasattr(est, "dataframe_index_in_")
    p1 = est.predict(Xte.tolist())
    est2 = pickle.loads(pickle.dumps(est))          # device handles are never pickled
    np.testing.assert_array_equal(est2.predict(Xte), p1)
    from sklearn.exceptions import NotFittedError

    with pytest.raises(NotFittedError, match="fitted with a dataframe"):
        est.kneighbors(return_dataframe_index=True)
    with pytest.raises(ValueError, match="n_neighbors <= n_samples_fit"):
        est.kneighbors(Xte, n_neighbors=1000)
    with pytest.raises(ValueError, match="features"):
        est.predict(Xte[:, :5])
    with pytest.raises(ValueError):
        est.predict(np.where(np.arange(Xte.size).reshape(Xte.shape) == 3, np.nan, Xte))
    gs = GridSearchCV(S.MSNRegressor(), {"n_neighbors": [1, 3]}, cv=2).fit(Xtr, ytr)
    assert gs.predict(Xte).shape == (33, ytr.shape[1])
    # 1-D target, float32 queries, F-ordered and read-only arrays
    e1 = S.RawKNNRegressor(n_neighbors=2).fit(Xtr, ytr[:, 0])
    assert e1.predict(Xte).shape == (33,)
    Xf = np.asfortranarray(Xte.astype(np.float32))
    Xf.setflags(write=False)
    assert est.predict(Xf).shape == (33, ytr.shape[1])
    # k = 1 on the training frame returns each row itself (ref:tests/test_estimators.py:181-201)
    idx = S.GNNRegressor(n_neighbors=1).fit(Xtr, ytr).kneighbors(Xtr, return_distance=False)
    np.testing.assert_array_equal(idx.ravel(), np.arange(len(Xtr)))


def test_integration_stub_from_the_document_runs():
    """The ctypes stub printed in INTEGRATION.md section 1 (what a maintainer of the reference would
    add) is executed verbatim against the built library and must agree with the package's own
    binding."""
    import re

    from sknnr_b200._build import lib_path
    from sknnr_b200._engine import KNNIndex
    from tests.conftest import ROOT
    import os

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(import ctypes as C, numpy as np.*?)```", text, re.S).group(1)
    code = code.replace('C.CDLL("libsknnr_b200.so")', f'C.CDLL({lib_path()!r})')
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    Xtr, Xte, ytr, yte, _ = _split()
    mean, scale = Xtr.mean(0), Xtr.std(0, ddof=1)
    fit_z = (Xtr - mean) / scale
    stub = ns["Index"](fit_z, ytr, mean, scale)
    d, i, p = stub.kneighbors(Xte, 5, transformed=False, deterministic=True, decimals=10, weights=1)
    d0, i0, p0 = KNNIndex(fit_z, mean, scale, None, ytr).query(Xte, 5, weights="uniform", with_pred=True)
    np.testing.assert_array_equal(i, i0)
    np.testing.assert_array_equal(d, d0)
    np.testing.assert_array_equal(p, p0)
    d, i, _ = stub.kneighbors(None, 5, transformed=True, deterministic=True, decimals=10)
    assert d.shape == (len(Xtr), 5) and not np.any(i == np.arange(len(Xtr))[:, None])
