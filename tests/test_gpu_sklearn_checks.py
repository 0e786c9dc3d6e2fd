"""scikit-learn's own estimator / transformer check suites run against this package (GPU), with the
same expected failures the reference declares for itself (ref:tests/test_estimators.py:64-134,
ref:tests/test_transformers.py:58-108): CCA-based classes cannot take the 1-D targets scikit-learn's
checks use, and the transformed estimators report the transformed feature count in
``n_features_in_``."""

import pytest
from sklearn.utils.estimator_checks import parametrize_with_checks

import sknnr_b200 as S
from sknnr_b200.transformers import (CCATransformer, CCorATransformer, GBNodeTransformer,
                                     MahalanobisTransformer, RFNodeTransformer, StandardScalerWithDOF)

pytestmark = [pytest.mark.gpu, pytest.mark.filterwarnings("ignore")]

_CCA_1D = [
    "check_estimators_dtypes", "check_dtype_object", "check_estimators_fit_returns_self",
    "check_pipeline_consistency", "check_estimators_overwrite_params", "check_fit_score_takes_y",
    "check_estimators_pickle", "check_methods_sample_order_invariance", "check_methods_subset_invariance",
    "check_dict_unchanged", "check_dont_overwrite_parameters", "check_fit_idempotent",
    "check_fit_check_is_fitted", "check_fit2d_predict1d", "check_fit2d_1sample", "check_estimators_nan_inf",
    "check_positive_only_tag_during_fit",
]
_GNN_ONLY = ["check_regressors_train", "check_regressor_data_not_an_array", "check_regressors_no_decision_function",
             "check_supervised_y_2d", "check_regressors_int"]
_GNN_ROW_SUMS = ["check_regressor_multioutput", "check_readonly_memmap_input", "check_f_contiguous_array_estimator"]
_CCA_TRANSFORMER_ONLY = ["check_transformer_data_not_an_array", "check_transformer_general",
                         "check_transformer_preserve_dtypes", "check_n_features_in", "check_requires_y_none",
                         "check_readonly_memmap_input", "check_n_features_in_after_fitting",
                         "check_f_contiguous_array_estimator"]


def _estimator_xfails(est):
    out = {}
    if isinstance(est, S.GNNRegressor):
        out.update({c: "CCA requires 2D y arrays." for c in _CCA_1D + _GNN_ONLY})
        out.update({c: "Row sums must be greater than 0." for c in _GNN_ROW_SUMS})
    if isinstance(est, (S.MSNRegressor, S.GNNRegressor, S.RFNNRegressor, S.GBNNRegressor)):
        out.update({c: "Estimator stores transformed n_features_in_"
                    for c in ("check_n_features_in_after_fitting", "check_n_features_in")})
    return out


def _transformer_xfails(t):
    if isinstance(t, CCATransformer):
        return {c: "CCA requires 2D y arrays." for c in _CCA_1D + _CCA_TRANSFORMER_ONLY}
    return {}


@parametrize_with_checks(
    [S.RawKNNRegressor(), S.EuclideanKNNRegressor(), S.MahalanobisKNNRegressor(), S.MSNRegressor(),
     S.GNNRegressor(), S.RFNNRegressor(), S.GBNNRegressor()],
    expected_failed_checks=_estimator_xfails)
def test_sklearn_estimator_checks(estimator, check):
    check(estimator)


@parametrize_with_checks(
    [StandardScalerWithDOF(), MahalanobisTransformer(), CCATransformer(), CCorATransformer(),
     GBNodeTransformer(), RFNodeTransformer()],
    expected_failed_checks=_transformer_xfails)
def test_sklearn_transformer_checks(estimator, check):
    check(estimator)
