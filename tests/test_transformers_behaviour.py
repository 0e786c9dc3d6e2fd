"""Behaviour of the transformers, after ref:tests/test_transformers.py, restated for this package.
The fit side (scikit-learn training, validation, weights, feature names) runs on the CPU; every
``transform`` goes through the device and is marked gpu."""

import warnings

import numpy as np
import pandas as pd
import pytest
from numpy.testing import assert_array_equal
from sklearn import config_context
from sklearn.datasets import make_classification
from sklearn.ensemble import (GradientBoostingClassifier, GradientBoostingRegressor,
                              RandomForestClassifier, RandomForestRegressor)
from sklearn.exceptions import NotFittedError
from sklearn.preprocessing import StandardScaler

from sknnr_b200.transformers import (CCATransformer, CCorATransformer, GBNodeTransformer,
                                     MahalanobisTransformer, RFNodeTransformer, StandardScalerWithDOF)
from tests.conftest import load_golden

pytestmark = pytest.mark.filterwarnings("ignore::FutureWarning")

ALL = [StandardScalerWithDOF, MahalanobisTransformer, CCATransformer, CCorATransformer,
       GBNodeTransformer, RFNodeTransformer]
ORDINATION = [CCATransformer, CCorATransformer]
TREES = [GBNodeTransformer, RFNodeTransformer]
FOREST_TYPES = {GBNodeTransformer: (GradientBoostingRegressor, GradientBoostingClassifier),
                RFNodeTransformer: (RandomForestRegressor, RandomForestClassifier)}


def _small(cls, **kw):
    if cls in TREES:
        kw.setdefault("n_estimators", 8)
    return cls(**kw)


@pytest.fixture(scope="module")
def moscow():
    g = load_golden("c2_moscow_gnn_k5.npz")
    X, y = g["X"], g["y_targets"]
    X_df = pd.DataFrame(X, columns=[f"band_{i}" for i in range(X.shape[1])], index=g["index"])
    return X, y, X_df


@pytest.mark.parametrize("cls", ALL)
def test_transform_raises_notfitted(cls, moscow):
    with pytest.raises(NotFittedError):
        cls().transform(moscow[0])


@pytest.mark.parametrize("x_type", ["array", "dataframe"])
@pytest.mark.parametrize("cls", ALL)
def test_feature_names_in_consistency(cls, x_type, moscow):
    X, y, X_df = moscow
    Xa = X_df if x_type == "dataframe" else X
    t = _small(cls).fit(Xa, y[:, :4] + 0.1)
    ref = StandardScaler().fit(Xa, y)
    if hasattr(ref, "feature_names_in_"):
        assert_array_equal(t.feature_names_in_, ref.feature_names_in_)
    else:
        assert not hasattr(t, "feature_names_in_")


@pytest.mark.parametrize("n_components", [-1, 1000])
@pytest.mark.parametrize("cls", ORDINATION)
def test_out_of_range_n_components(cls, n_components, moscow):
    with pytest.raises(ValueError, match=r"n_components=-?\d+ must be between 0 and \d+"):
        cls(n_components=n_components).fit(moscow[0], moscow[1])


@pytest.mark.parametrize("cls", TREES)
def test_forest_types_follow_the_target_dtype(cls, moscow):
    X, y, _ = moscow
    reg, clf = FOREST_TYPES[cls]
    est = _small(cls).fit(X, y[:, :3])
    assert set(est.estimator_type_dict_.values()) == {"regression"}
    assert all(isinstance(f, reg) for f in est.estimators_)
    yb = y[:, :3].astype(bool)
    yb[0] = ~yb[0]                      # boosting needs two classes in every target
    est = _small(cls).fit(X, yb)
    assert set(est.estimator_type_dict_.values()) == {"classification"}
    assert all(isinstance(f, clf) for f in est.estimators_)


@pytest.mark.parametrize("nan_like", [np.nan, None, pd.NA])
@pytest.mark.parametrize("wrap", [pd.Series, np.asarray])
@pytest.mark.parametrize("cls", TREES)
def test_nan_like_targets_raise(cls, wrap, nan_like, moscow):
    X, y, _ = moscow
    t = y[:, 0].astype(object)
    t[0] = nan_like
    with pytest.raises(ValueError, match=r"Target \S+ has NaN-like elements"):
        _small(cls).fit(X, wrap(t, dtype=object))


@pytest.mark.parametrize("wrap", [pd.Series, np.asarray])
@pytest.mark.parametrize("cls", TREES)
def test_mixed_string_targets_raise(cls, wrap, moscow):
    X, y, _ = moscow
    t = y[:, 0].astype(object)
    t[-1] = "mixed"
    with pytest.raises(ValueError, match=r"Target \S+ has non-string types"):
        _small(cls).fit(X, wrap(t, dtype=object))


@pytest.mark.parametrize("cls", TREES)
def test_duplicate_target_names(cls):
    X = pd.DataFrame(np.random.default_rng(0).random((2, 2)), columns=["f1", "f2"])
    with pytest.raises(ValueError, match=r"Duplicate feature names found: \['a'\]\.$"):
        _small(cls).fit(X, pd.DataFrame([[1, 2], [3, 4]], columns=["a", "a"]))
    t = _small(cls).fit(X, pd.DataFrame([[1, 2], [3, 4]], columns=[1, "1"]))
    assert t.estimator_type_dict_ == {1: "regression", "1": "regression"}


@pytest.mark.parametrize("max_features_reg", ["sqrt", "log2"])
@pytest.mark.parametrize("max_features_clf", ["log2", 1.0])
def test_rfnode_non_default_parameters_reach_the_right_forests(max_features_reg, max_features_clf, moscow):
    X, y, _ = moscow
    yd = pd.DataFrame(y[:, :3], columns=["a", "b", "c"])
    yd["present"] = yd["a"] > 0.0
    est = RFNodeTransformer(n_estimators=5, criterion_reg="absolute_error", criterion_clf="entropy",
                            max_features_reg=max_features_reg, max_features_clf=max_features_clf,
                            class_weight_clf="balanced_subsample").fit(X, yd)
    kinds = set(est.estimator_type_dict_.values())
    assert kinds == {"regression", "classification"}
    for rf in est.estimators_:
        p = rf.get_params()
        if isinstance(rf, RandomForestClassifier):
            assert (p["criterion"], p["max_features"], p["class_weight"]) == ("entropy", max_features_clf, "balanced_subsample")
        else:
            assert (p["criterion"], p["max_features"]) == ("absolute_error", max_features_reg)


def test_gbnode_non_default_parameters_reach_the_right_models(moscow):
    X, y, _ = moscow
    yd = pd.DataFrame(y[:, :2], columns=["a", "b"])
    yd["present"] = yd["a"] > 0.0
    est = GBNodeTransformer(n_estimators=5, loss_reg="absolute_error", loss_clf="exponential", alpha_reg=0.1).fit(X, yd)
    assert set(est.estimator_type_dict_.values()) == {"regression", "classification"}
    for gb in est.estimators_:
        p = gb.get_params()
        if isinstance(gb, GradientBoostingClassifier):
            assert p["loss"] == "exponential"
        else:
            assert (p["loss"], p["alpha"]) == ("absolute_error", 0.1)


@pytest.mark.parametrize("method", ["train_improvement", "uniform"])
@pytest.mark.parametrize("n_classes", [2, 3, 5])
def test_gbnode_multiclass_attributes(method, n_classes):
    X, yc = make_classification(n_samples=100, n_features=20, n_informative=10, n_classes=n_classes, random_state=42)
    y = np.array([yc.astype(str), yc % 2, yc.astype(float)], dtype=object).T
    est = GBNodeTransformer(n_estimators=12, tree_weighting_method=method).fit(X, y)
    per_iter = 1 if n_classes == 2 else n_classes
    assert est.n_forests_ == 3 and est.n_trees_per_iteration_ == [per_iter, 1, 1]
    assert [w.shape for w in est.tree_weights_] == [(12 * per_iter,), (12,), (12,)]
    assert len(est._trees()) == 12 * (per_iter + 2)
    assert len(est.get_feature_names_out()) == 12 * (per_iter + 2)


# ---- everything below calls transform: device --------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cls", ALL)
def test_feature_names_out_count_matches_transform(cls, moscow):
    X, y, _ = moscow
    t = _small(cls).fit(X, y[:, :5] + 0.1)
    assert t.get_feature_names_out().shape == (t.transform(X).shape[1],)


@pytest.mark.gpu
@pytest.mark.parametrize("config_type", ["global", "transformer"])
@pytest.mark.parametrize("output_mode", ["default", "pandas"])
@pytest.mark.parametrize("x_type", ["array", "dataframe"])
@pytest.mark.parametrize("cls", ALL)
def test_transform_output_type_consistency(cls, x_type, output_mode, config_type, moscow):
    X, y, X_df = moscow
    Xa = X_df if x_type == "dataframe" else X
    t, ref = _small(cls), StandardScaler()
    cfg = {}
    if config_type == "global":
        cfg = {"transform_output": output_mode}
    else:
        t.set_output(transform=output_mode)
        ref.set_output(transform=output_mode)
    with config_context(**cfg):
        ours = type(t.fit(Xa, y[:, :4] + 0.1).transform(Xa))
        theirs = type(ref.fit(Xa, y).transform(Xa))
    assert ours is theirs


@pytest.mark.gpu
@pytest.mark.parametrize("n_components", [None, 0, 5])
@pytest.mark.parametrize("cls", ORDINATION)
def test_n_components(cls, n_components, moscow):
    X, y, _ = moscow
    t = cls(n_components=n_components).fit(X, y)
    if n_components is not None:
        assert t.n_components_ == n_components
    assert t.transform(X).shape[1] == t.n_components_


@pytest.mark.gpu
@pytest.mark.parametrize("n_classes", [2, 3])
def test_gbnode_multiclass_transform_shape_and_order(n_classes):
    X, yc = make_classification(n_samples=100, n_features=20, n_informative=10, n_classes=n_classes, random_state=42)
    y = np.array([yc.astype(str), yc % 2, yc.astype(float)], dtype=object).T
    est = GBNodeTransformer(n_estimators=9).fit(X, y)
    ids = est.transform(X)
    assert ids.shape == (100, sum(est.n_trees_per_iteration_) * 9) and ids.dtype == np.int64
    want = []
    for e in est.estimators_:
        a = e.apply(X)
        want.append(np.swapaxes(a, 1, 2).reshape(len(X), -1) if a.ndim == 3 else a)
    assert_array_equal(ids, np.hstack(want).astype(np.int64))
