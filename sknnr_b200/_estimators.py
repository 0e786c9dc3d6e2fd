"""The estimators of the hot-path scope: one-method subclasses that choose the transformer
(mirrors ref:src/sknnr/_euclidean.py:55-56, _mahalanobis.py:56-57, _msn.py:68-69, _gnn.py:69-70,
_weighted_trees.py:18-140, _rfnn.py:16-239 and - scope row f3 - _gbnn.py:16-249)."""

from __future__ import annotations

import numpy as np

from ._base import OrdinationKNeighborsRegressor, TransformedKNeighborsRegressor, YFitMixin
from .transformers import (
    CCATransformer,
    CCorATransformer,
    GBNodeTransformer,
    MahalanobisTransformer,
    RFNodeTransformer,
    StandardScalerWithDOF,
)


class EuclideanKNNRegressor(TransformedKNeighborsRegressor):
    """kNN regression in standardised (N-1 dof) feature space."""

    def _get_transformer(self):
        return StandardScalerWithDOF(ddof=1)


class MahalanobisKNNRegressor(TransformedKNeighborsRegressor):
    """kNN regression in Mahalanobis (whitened) feature space."""

    def _get_transformer(self):
        return MahalanobisTransformer()


class MSNRegressor(YFitMixin, OrdinationKNeighborsRegressor):
    """Most Similar Neighbour imputation: canonical correlation (CCorA) space."""

    def _get_transformer(self):
        return CCorATransformer(self.n_components)


class GNNRegressor(YFitMixin, OrdinationKNeighborsRegressor):
    """Gradient Nearest Neighbour imputation: canonical correspondence (CCA) space."""

    def _get_transformer(self):
        return CCATransformer(self.n_components)


class WeightedTreesNNRegressor(YFitMixin, TransformedKNeighborsRegressor):
    """Tree-based transformed regressors searching with the weighted Hamming metric over node
    IDs (mirrors ref:src/sknnr/_weighted_trees.py)."""

    def __init__(self, *, n_neighbors=5, weights="uniform", n_jobs=None):
        super().__init__(n_neighbors=n_neighbors, weights=weights, algorithm="brute",
                         metric="hamming", n_jobs=n_jobs)

    def _set_fitted_transformer(self, X, y) -> None:
        super()._set_fitted_transformer(X, y)
        self.hamming_weights_ = self._get_hamming_weights()

    def _get_hamming_weights(self):
        n_forests = self.transformer_.n_forests_
        if isinstance(self.forest_weights, str) and self.forest_weights == "uniform":
            fw = np.full(n_forests, 1.0 / n_forests, dtype=np.float64)
        else:
            fw = self._validate_user_forest_weights()
            fw /= np.sum(fw)
        for i in range(len(fw)):
            fw[i] /= self.transformer_.n_trees_per_iteration_[i]
        return np.hstack([tw * f for tw, f in zip(self.transformer_.tree_weights_, fw, strict=True)])

    def _validate_user_forest_weights(self):
        n_forests = self.transformer_.n_forests_
        try:
            fw = np.asarray(self.forest_weights, dtype=np.float64)
        except (TypeError, ValueError) as e:
            raise ValueError(
                f"`forest_weights` must be a sequence of numeric values, "
                f"but got {self.forest_weights} instead.") from e
        if fw.shape != (n_forests,):
            raise ValueError(f"Expected `forest_weights` to have length {n_forests}, but got {fw.size}.")
        if not np.all(np.isfinite(fw)):
            raise ValueError(f"Expected elements in `forest_weights` to be finite, but got {fw}.")
        if np.any(fw < 0):
            raise ValueError(f"Expected elements in `forest_weights` to be non-negative, but got {fw}.")
        if np.sum(fw) <= 0:
            raise ValueError(f"At least one element in `forest_weights` must be positive, but got {fw}.")
        return fw

    def _get_additional_regressor_init_kwargs(self) -> dict:
        return {"metric_params": {"w": self.hamming_weights_}}


class RFNNRegressor(WeightedTreesNNRegressor):
    """Random Forest Nearest Neighbours imputation (mirrors ref:src/sknnr/_rfnn.py:16-239)."""

    def __init__(self, *, n_estimators=50, criterion_reg="squared_error", criterion_clf="gini",
                 max_depth=None, min_samples_split=2, min_samples_leaf=5,
                 min_weight_fraction_leaf=0.0, max_features_reg=1.0, max_features_clf="sqrt",
                 max_leaf_nodes=None, min_impurity_decrease=0.0, bootstrap=True, oob_score=False,
                 n_jobs=None, random_state=None, verbose=0, warm_start=False,
                 class_weight_clf=None, ccp_alpha=0.0, max_samples=None, monotonic_cst=None,
                 forest_weights="uniform", n_neighbors=5, weights="uniform"):
        self.n_estimators = n_estimators
        self.criterion_reg = criterion_reg
        self.criterion_clf = criterion_clf
        self.max_depth = max_depth
        self.min_samples_split = min_samples_split
        self.min_samples_leaf = min_samples_leaf
        self.min_weight_fraction_leaf = min_weight_fraction_leaf
        self.max_features_reg = max_features_reg
        self.max_features_clf = max_features_clf
        self.max_leaf_nodes = max_leaf_nodes
        self.min_impurity_decrease = min_impurity_decrease
        self.bootstrap = bootstrap
        self.oob_score = oob_score
        self.n_jobs = n_jobs
        self.random_state = random_state
        self.verbose = verbose
        self.warm_start = warm_start
        self.class_weight_clf = class_weight_clf
        self.ccp_alpha = ccp_alpha
        self.max_samples = max_samples
        self.monotonic_cst = monotonic_cst
        self.forest_weights = forest_weights
        super().__init__(n_neighbors=n_neighbors, weights=weights, n_jobs=self.n_jobs)

    def _get_transformer(self):
        names = ["n_estimators", "criterion_reg", "criterion_clf", "max_depth", "min_samples_split",
                 "min_samples_leaf", "min_weight_fraction_leaf", "max_features_reg",
                 "max_features_clf", "max_leaf_nodes", "min_impurity_decrease", "bootstrap",
                 "oob_score", "n_jobs", "random_state", "verbose", "warm_start", "class_weight_clf",
                 "ccp_alpha", "max_samples", "monotonic_cst"]
        return RFNodeTransformer(**{n: getattr(self, n) for n in names})


class GBNNRegressor(WeightedTreesNNRegressor):
    """Gradient Boosting Nearest Neighbours imputation (mirrors ref:src/sknnr/_gbnn.py:16-249):
    one boosted model per target, neighbours by Hamming distance over node IDs with every tree
    weighted by its share of the training-loss reduction."""

    def __init__(self, *, loss_reg="squared_error", loss_clf="log_loss", learning_rate=0.1,
                 n_estimators=100, subsample=1.0, criterion="friedman_mse", min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_depth=3,
                 min_impurity_decrease=0.0, init=None, random_state=None, max_features=None,
                 alpha_reg=0.9, verbose=0, max_leaf_nodes=None, warm_start=False,
                 validation_fraction=0.1, n_iter_no_change=None, tol=0.0001, ccp_alpha=0.0,
                 forest_weights="uniform", tree_weighting_method="train_improvement",
                 n_neighbors=5, weights="uniform", n_jobs=None):
        self.loss_reg = loss_reg
        self.loss_clf = loss_clf
        self.learning_rate = learning_rate
        self.n_estimators = n_estimators
        self.subsample = subsample
        self.criterion = criterion
        self.min_samples_split = min_samples_split
        self.min_samples_leaf = min_samples_leaf
        self.min_weight_fraction_leaf = min_weight_fraction_leaf
        self.max_depth = max_depth
        self.min_impurity_decrease = min_impurity_decrease
        self.init = init
        self.random_state = random_state
        self.max_features = max_features
        self.alpha_reg = alpha_reg
        self.verbose = verbose
        self.max_leaf_nodes = max_leaf_nodes
        self.warm_start = warm_start
        self.validation_fraction = validation_fraction
        self.n_iter_no_change = n_iter_no_change
        self.tol = tol
        self.ccp_alpha = ccp_alpha
        self.forest_weights = forest_weights
        self.tree_weighting_method = tree_weighting_method
        super().__init__(n_neighbors=n_neighbors, weights=weights, n_jobs=n_jobs)

    def _get_transformer(self):
        names = ["loss_reg", "loss_clf", "learning_rate", "n_estimators", "subsample", "criterion",
                 "min_samples_split", "min_samples_leaf", "min_weight_fraction_leaf", "max_depth",
                 "min_impurity_decrease", "init", "random_state", "max_features", "alpha_reg",
                 "verbose", "max_leaf_nodes", "warm_start", "validation_fraction",
                 "n_iter_no_change", "tol", "ccp_alpha", "tree_weighting_method"]
        return GBNodeTransformer(**{n: getattr(self, n) for n in names})
