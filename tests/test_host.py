"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol,
fit-side transformer math reproduces the reference's fitted state, node-code mapping, and the
estimator surface's error behaviour that must not need a device."""

import ctypes
import inspect
import os
import re

import numpy as np
import pytest
from sklearn.exceptions import NotFittedError

from tests.conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from sknnr_b200 import _lib as L

    header = open(os.path.join(ROOT, "include", "sknnr_b200.h")).read()
    declared = set(re.findall(r"\b(sknnr_[a-z0-9_]+)\s*\(", header))
    declared -= {"sknnr_index", "sknnr_stats", "sknnr_hamming_index"}
    assert declared == set(L.EXPORTS)
    lib = ctypes.CDLL(L.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert L.load().sknnr_abi_version() == L.ABI_VERSION
    assert ctypes.sizeof(L.Stats) == 56


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry point fails loudly."""
    from sknnr_b200 import _lib as L
    from sknnr_b200._engine import HammingIndex, KNNIndex

    if L.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(L.SknnrError, match="no CUDA device"):
        KNNIndex(np.zeros((4, 2)))
    with pytest.raises(L.SknnrError, match="no CUDA device"):
        HammingIndex(np.zeros((4, 2), dtype=np.uint16), np.ones(2))


def test_product_never_imports_the_oracle():
    import sknnr_b200

    pkg = os.path.dirname(sknnr_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("sknnr_oracle", "oracle") or f == "__never__", (dirpath, f)


@pytest.mark.parametrize(("name", "cls_name", "kw"), [
    ("euclidean", "StandardScalerWithDOF", {"ddof": 1}),
    ("mahalanobis", "MahalanobisTransformer", {}),
    ("msn", "CCorATransformer", {}),
    ("gnn", "CCATransformer", {}),
])
@pytest.mark.parametrize("comp", ["full", "reduced"])
def test_fit_side_state_matches_reference(name, cls_name, kw, comp):
    """Fit math (NumPy, cold path) must hand the device the same (centre, scale, projector) the
    reference's transformers hold (ref:src/sknnr/transformers/*.py fit methods)."""
    import sknnr_b200.transformers as T

    if comp == "reduced":
        if name not in ("msn", "gnn"):
            pytest.skip("no n_components")
        kw = dict(kw, n_components=3)
    g = load_golden(f"moscow_{name}_{comp}.npz")
    sp = load_golden("moscow_split.npz")
    t = getattr(T, cls_name)(**kw).fit(sp["X_train"], sp["y_train"])
    center, scale, proj, d_out = t._affine()
    for key, arr in (("state_center", center), ("state_scale", scale), ("state_proj", proj)):
        if arr is None:
            assert key not in g
        else:
            np.testing.assert_allclose(arr, g[key], rtol=1e-10, atol=1e-12)
    assert d_out == g["state_fit_Z"].shape[1]


def test_n_components_validation():
    from sknnr_b200.transformers import CCATransformer, CCorATransformer

    sp = load_golden("moscow_split.npz")
    for cls in (CCATransformer, CCorATransformer):
        with pytest.raises(ValueError, match="n_components=99 must be between"):
            cls(n_components=99).fit(sp["X_train"], sp["y_train"])
        t = cls(n_components=2).fit(sp["X_train"], sp["y_train"])
        assert t.n_components_ == 2 and len(t.get_feature_names_out()) == 2


def test_node_code_mapping():
    from sknnr_b200._base import _encode_nodes, _node_code_tables

    ref = np.array([[3, 100000], [7, 5], [3, 70000]])
    tables = _node_code_tables(ref)
    assert tables is not None
    enc = _encode_nodes(ref, tables)
    assert enc.dtype == np.uint16 and enc[0, 0] == enc[2, 0] != enc[1, 0]
    q = _encode_nodes(np.array([[7, 123456], [4, 5]]), tables)
    assert q[0, 0] == enc[1, 0] and q[0, 1] == 31743 and q[1, 0] == 31743 and q[1, 1] == enc[1, 1]
    small = np.array([[0, 5], [9, 2]])
    assert _node_code_tables(small) is None
    assert _encode_nodes(np.array([[40000, -1]]), None).tolist() == [[31743, 31743]]


def test_signatures_match_reference_surface():
    """Constructor signatures of ref:src/sknnr/_base.py:196-218,385-408 and _rfnn.py:154-214."""
    import sknnr_b200 as S

    base = ["n_neighbors", "weights", "algorithm", "leaf_size", "p", "metric", "metric_params", "n_jobs"]
    for cls in (S.EuclideanKNNRegressor, S.MahalanobisKNNRegressor):
        assert list(inspect.signature(cls).parameters) == base
    for cls in (S.MSNRegressor, S.GNNRegressor):
        assert list(inspect.signature(cls).parameters) == ["n_neighbors", "n_components"] + base[1:]
    rf = list(inspect.signature(S.RFNNRegressor).parameters)
    assert rf[0] == "n_estimators" and rf[-3:] == ["forest_weights", "n_neighbors", "weights"] and len(rf) == 24
    kn = list(inspect.signature(S.RawKNNRegressor.kneighbors).parameters)
    assert kn == ["self", "X", "n_neighbors", "return_distance", "return_dataframe_index",
                  "use_deterministic_ordering"]
    assert S.RawKNNRegressor.DISTANCE_PRECISION_DECIMALS == 10
    for cls in (S.EuclideanKNNRegressor, S.MSNRegressor, S.RFNNRegressor, S.RawKNNRegressor):
        est = cls()
        assert type(est)(**est.get_params()).get_params().keys() == est.get_params().keys()


@pytest.mark.parametrize("name", ["RawKNNRegressor", "EuclideanKNNRegressor", "MahalanobisKNNRegressor",
                                  "MSNRegressor", "GNNRegressor", "RFNNRegressor"])
def test_unfitted_estimators_raise(name):
    import sknnr_b200 as S

    X = np.zeros((3, 4))
    with pytest.raises(NotFittedError):
        getattr(S, name)().kneighbors(X)
    with pytest.raises(NotFittedError):
        getattr(S, name)().predict(X)
