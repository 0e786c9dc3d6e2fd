"""Behavioural surface of the reference's estimators, restated for this package (GPU): the cases
of ref:tests/test_estimators.py - NotFitted errors, lists / dataframes / dataframe indexes,
feature-name warnings, output types, y_fit, GridSearchCV, n_features_in_, deterministic ordering
and its precision, forest weights - run against sknnr_b200 so that a user of the reference finds the
same behaviour.  Data: the Moscow Mountain / St. Joe's plots shipped in tests/golden."""

import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_array_equal
from sklearn import config_context
from sklearn.exceptions import NotFittedError
from sklearn.model_selection import GridSearchCV
from sklearn.neighbors import KNeighborsRegressor

import sknnr_b200 as S
from tests.conftest import load_golden

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore::FutureWarning")]

ALL = ["RawKNNRegressor", "EuclideanKNNRegressor", "MahalanobisKNNRegressor", "MSNRegressor",
       "GNNRegressor", "RFNNRegressor", "GBNNRegressor"]
TRANSFORMED = ALL[1:]
YFIT = ["MSNRegressor", "GNNRegressor", "RFNNRegressor", "GBNNRegressor"]
TREES = ["RFNNRegressor", "GBNNRegressor"]


def _make(name, **kw):
    # boosted models are the slow part of these tests: fewer stages where the count does not matter
    if name == "GBNNRegressor":
        kw.setdefault("n_estimators", 20)
    if name == "RFNNRegressor":
        kw.setdefault("n_estimators", 20)
    return getattr(S, name)(**kw)


@pytest.fixture(scope="module")
def moscow():
    g = load_golden("c2_moscow_gnn_k5.npz")
    X, y, index = g["X"], g["y_targets"], g["index"]
    cols = [f"band_{i}" for i in range(X.shape[1])]
    X_df = pd.DataFrame(X, columns=cols, index=index)
    y_df = pd.DataFrame(y, columns=[f"sp_{i}" for i in range(y.shape[1])], index=index)
    return X, y, index, X_df, y_df


@pytest.fixture(scope="module")
def X_y_yfit(moscow):
    X, y = moscow[0], moscow[1]
    return X, y[:, :10] + 0.1, y[:, 10:] + 0.1


@pytest.mark.parametrize("name", ALL)
def test_continuous_multioutput_lists_and_dataframes(name, moscow):
    X, y, _, X_df, y_df = moscow
    y = y[:, :6] + 0.1        # (a constant keeps every row sum positive, which CCA needs)
    p0 = _make(name, random_state=0).fit(X, y).predict(X) if name in TREES else _make(name).fit(X, y).predict(X)
    assert p0.shape == y.shape
    kw = {"random_state": 0} if name in TREES else {}
    p1 = _make(name, **kw).fit(X.tolist(), y.tolist()).predict(X.tolist())
    assert_array_equal(p1, p0)
    p2 = _make(name, **kw).fit(X_df, y_df.iloc[:, :6] + 0.1).predict(X_df)
    assert_array_equal(p2, p0)


@pytest.mark.parametrize("name", ALL)
def test_dataframe_indexes(name, moscow):
    X, y, index, X_df, _ = moscow
    y = y[:, :6] + 0.1        # (a constant keeps every row sum positive, which CCA needs)
    # (the tree estimators keep their default ensemble sizes here: with few trees two plots can share
    # every leaf, and a plot tied with another at distance 0 need not be its own first neighbour)
    est = getattr(S, name)(n_neighbors=1)
    est.fit(X, y)
    with pytest.raises(NotFittedError, match="fitted with a dataframe"):
        est.kneighbors(return_dataframe_index=True)
    est.fit(X.tolist(), y)                       # list.index must not be mistaken for an index
    assert not hasattr(est, "dataframe_index_in_")
    est.fit(X_df, y)
    assert_array_equal(est.dataframe_index_in_, index)
    idx = est.kneighbors(X_df, return_distance=False, return_dataframe_index=True)
    assert_array_equal(idx.ravel(), index)       # k = 1: every plot finds itself


@pytest.mark.parametrize("fit_names", [True, False])
@pytest.mark.parametrize("name", ALL)
def test_warn_for_missing_feature_names(name, fit_names, moscow):
    X, y, _, X_df, _ = moscow
    y = y[:, :6] + 0.1        # (a constant keeps every row sum positive, which CCA needs)
    msg = "fitted with feature names" if fit_names else "fitted without feature names"
    fit_X, predict_X = (X_df, X) if fit_names else (X, X_df)
    est = _make(name).fit(fit_X, y)
    with pytest.warns(UserWarning, match=msg):
        est.predict(predict_X)


@pytest.mark.parametrize("output_mode", ["default", "pandas"])
@pytest.mark.parametrize("as_frame", [False, True])
@pytest.mark.parametrize("name", ALL)
def test_output_type_consistency(name, as_frame, output_mode, moscow):
    X, y, _, X_df, y_df = moscow
    Xa, ya = (X_df, y_df.iloc[:, :6] + 0.1) if as_frame else (X, y[:, :6] + 0.1)
    with config_context(transform_output=output_mode):
        ours = type(_make(name).fit(Xa, ya).predict(Xa))
        ref = type(KNeighborsRegressor().fit(Xa, ya).predict(Xa))
    assert ours is ref


@pytest.mark.parametrize("name", YFIT)
def test_yfit_is_stored_and_affects_prediction(name, X_y_yfit):
    X, y, y_fit = X_y_yfit
    kw = {"random_state": 0} if name in TREES else {}
    est = _make(name, **kw).fit(X, y)
    assert est.y_fit_ is None
    without = est.independent_prediction_
    est.fit(X, y, y_fit=y_fit)
    assert_array_equal(est.y_fit_, y_fit)
    assert not np.array_equal(est.independent_prediction_, without)


@pytest.mark.parametrize("name", ALL)
def test_gridsearchcv(name, X_y_yfit):
    X, y, _ = X_y_yfit
    gs = GridSearchCV(_make(name), param_grid={"n_neighbors": [1, 3]}, cv=2, error_score="raise")
    gs.fit(X, y)
    assert gs.predict(X).shape == y.shape
    assert np.all(np.isfinite(gs.cv_results_["mean_test_score"]))


@pytest.mark.parametrize("name", TRANSFORMED)
def test_n_features_in(name, X_y_yfit):
    X, y, _ = X_y_yfit
    est = _make(name).fit(X, y)
    assert est.transformer_.n_features_in_ == X.shape[1]
    assert est.n_features_in_ == len(est.transformer_.get_feature_names_out())


@pytest.mark.parametrize(("deterministic", "expected"), [(False, [1, 0]), (True, [0, 1])])
def test_kneighbors_deterministic_ordering(deterministic, expected):
    X = np.array([1e-11, 1e-12, 1.0]).reshape(-1, 1)
    _, idx = S.RawKNNRegressor(n_neighbors=2).fit(X, np.array([0, 1, 2])).kneighbors(
        np.array([[0.0]]), use_deterministic_ordering=deterministic)
    assert_array_equal(idx[0], expected)


def test_kneighbors_uses_index_difference():
    X = np.array([1e-11, 1e-12, 1.0]).reshape(-1, 1)
    _, idx = S.RawKNNRegressor(n_neighbors=2).fit(X, np.array([0, 1, 2])).kneighbors(
        np.array([[0.0], [0.0]]), use_deterministic_ordering=True)
    assert_array_equal(idx[0], [0, 1])
    assert_array_equal(idx[1], [1, 0])


@pytest.mark.parametrize(("decimals", "expected"), [(8, [2, 1, 0]), (5, [1, 2, 0]), (2, [0, 1, 2])])
def test_kneighbors_precision_decimals(monkeypatch, decimals, expected):
    monkeypatch.setattr(S.RawKNNRegressor, "DISTANCE_PRECISION_DECIMALS", decimals)
    X = np.array([1e-3, 1e-6, 1e-9, 1.0]).reshape(-1, 1)
    _, idx = S.RawKNNRegressor(n_neighbors=3).fit(X, np.array([0, 1, 2, 3])).kneighbors(
        np.array([[0.0]]), use_deterministic_ordering=True)
    assert_array_equal(idx[0], expected)


def _std_weights(fw, n):
    if isinstance(fw, str):
        return np.full(n, 1.0 / n)
    a = np.asarray(fw, dtype=np.float64)
    return a / a.sum()


@pytest.mark.parametrize("forest_weights", ["uniform", [0.5, 1.5], (1.0, 2.0), np.array([3.0, 1.0])])
@pytest.mark.parametrize("name", TREES)
def test_tree_estimator_forest_weights(name, forest_weights, moscow):
    X, y = moscow[0], moscow[1][:, :2]
    est = _make(name, forest_weights=forest_weights, random_state=0).fit(X, y)
    got = est.hamming_weights_.reshape(est.transformer_.n_forests_, -1).sum(axis=1)
    np.testing.assert_allclose(got, _std_weights(forest_weights, 2), atol=1e-3)
    assert est.hamming_weights_.sum() == pytest.approx(1.0)
    assert est.predict(X).shape == y.shape


@pytest.mark.parametrize("forest_weights", ["uniform", [0.5, 1.5]])
def test_gbnn_multiclass_weights(forest_weights, moscow):
    X, y = moscow[0], moscow[1]
    cls = np.digitize(y[:, 0], np.percentile(y[:, 0], [33, 66])).astype(str)
    y_fit = pd.DataFrame({"total": y[:, 1], "cls": cls})
    est = S.GBNNRegressor(n_estimators=15, forest_weights=forest_weights, random_state=0).fit(X, y[:, :4], y_fit=y_fit)
    tr = est.transformer_
    assert tr.n_trees_per_iteration_ == [1, 3]
    assert est.hamming_weights_.shape == (15 + 45,)
    per_group = est.hamming_weights_.reshape(-1, 15).sum(axis=1)       # [reg, cls0, cls1, cls2]
    fw = _std_weights(forest_weights, 2)
    np.testing.assert_allclose(per_group, [fw[0], fw[1] / 3, fw[1] / 3, fw[1] / 3], atol=1e-3)
    assert est.kneighbors(X[:5])[1].shape == (5, 5)


@pytest.mark.parametrize(("bad", "msg"), [
    ([1.0], "to have length"), ([1.0, np.nan], "finite"), ([1.0, -1.0], "non-negative"),
    ([0.0, 0.0], "must be positive"), (["a", "b"], "numeric"),
])
@pytest.mark.parametrize("name", TREES)
def test_tree_estimator_raises_on_invalid_forest_weights(name, bad, msg, moscow):
    X, y = moscow[0], moscow[1][:, :2]
    with pytest.raises(ValueError, match=msg):
        _make(name, forest_weights=bad, n_estimators=3).fit(X, y)
