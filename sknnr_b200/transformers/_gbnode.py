"""Gradient-boosting node transformer for GBNN (mirrors
ref:src/sknnr/transformers/_gbnode_transformer.py).  One ``GradientBoostingRegressor`` /
``GradientBoostingClassifier`` per target is TRAINED by scikit-learn (cold path); ``transform``
walks all boosted trees on the GPU.  Multiclass classifiers contribute ``n_classes`` trees per
boosting stage; ``transform`` orders their columns class-major, stage-minor like the reference
(ref:src/sknnr/transformers/_tree_node_transformer.py:190-200).
"""

from __future__ import annotations

import numpy as np
from sklearn._loss.loss import HalfBinomialLoss, HalfSquaredError
from sklearn.ensemble import GradientBoostingClassifier, GradientBoostingRegressor
from sklearn.utils.validation import check_is_fitted

from ._treenode import TreeNodeTransformer

_GB_COMMON = ("learning_rate", "n_estimators", "subsample", "criterion", "min_samples_split",
              "min_samples_leaf", "min_weight_fraction_leaf", "max_depth", "min_impurity_decrease",
              "init", "random_state", "max_features", "verbose", "max_leaf_nodes", "warm_start",
              "validation_fraction", "n_iter_no_change", "tol", "ccp_alpha")


def train_improvement(est, X, y):
    """Per-stage share of the training-loss reduction of a fitted gradient-boosting model
    (ref:src/sknnr/transformers/_gbnode_transformer.py:20-56): the loss before stage 0 is the
    loss of the init prediction (doubled for the "half" losses, as scikit-learn reports
    ``train_score_``), every stage's delta is divided by the total; all-zero deltas give ones."""
    if hasattr(est, "classes_"):
        y = np.searchsorted(est.classes_, y).astype("float64")
    factor = 2 if isinstance(est._loss, (HalfSquaredError, HalfBinomialLoss)) else 1
    start = est._loss(np.asarray(y, dtype=np.float64), est._raw_predict_init(X)) * factor
    delta = np.diff(np.hstack([start, est.train_score_]))
    if np.allclose(delta, 0.0):
        return np.ones_like(delta, dtype=np.float64)
    return delta / np.sum(delta)


class GBNodeTransformer(TreeNodeTransformer):
    def __init__(self, loss_reg="squared_error", loss_clf="log_loss", learning_rate=0.1,
                 n_estimators=100, subsample=1.0, criterion="friedman_mse", min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_depth=3,
                 min_impurity_decrease=0.0, init=None, random_state=None, max_features=None,
                 alpha_reg=0.9, verbose=0, max_leaf_nodes=None, warm_start=False,
                 validation_fraction=0.1, n_iter_no_change=None, tol=0.0001, ccp_alpha=0.0,
                 tree_weighting_method="train_improvement"):
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
        self.alpha_reg = alpha_reg
        self.init = init
        self.random_state = random_state
        self.max_features = max_features
        self.verbose = verbose
        self.max_leaf_nodes = max_leaf_nodes
        self.warm_start = warm_start
        self.validation_fraction = validation_fraction
        self.n_iter_no_change = n_iter_no_change
        self.tol = tol
        self.ccp_alpha = ccp_alpha
        self.tree_weighting_method = tree_weighting_method

    def fit(self, X, y):
        common = {name: getattr(self, name) for name in _GB_COMMON}
        return self._fit(
            X, y,
            lambda: GradientBoostingRegressor(loss=self.loss_reg, alpha=self.alpha_reg, **common),
            lambda: GradientBoostingClassifier(loss=self.loss_clf, **common))

    def _set_n_trees_per_iteration(self):
        return [est.n_trees_per_iteration_ for est in self.estimators_]

    def _set_tree_weights(self, X, y):
        out = []
        if self.tree_weighting_method == "train_improvement":
            for est, target in zip(self.estimators_, y, strict=True):
                w = train_improvement(est, X, target)
                w /= w.sum()
                out.append(np.tile(w, est.n_trees_per_iteration_))
        elif self.tree_weighting_method == "uniform":
            for est in self.estimators_:
                n = est.n_estimators * est.n_trees_per_iteration_
                out.append(np.full(n, 1.0 / n, dtype=np.float64))
        else:
            raise ValueError(
                f"Invalid tree_weighting_method: {self.tree_weighting_method}. "
                "Must be 'train_improvement' or 'uniform'.")
        return out

    def _trees(self):
        # est.estimators_ is [n_stages, n_trees_per_iteration]; columns run class-major
        return [est.estimators_[s, c].tree_ for est in self.estimators_
                for c in range(est.estimators_.shape[1]) for s in range(est.estimators_.shape[0])]

    def get_feature_names_out(self, input_features=None):
        check_is_fitted(self, "estimators_")
        names = []
        for i, est in enumerate(self.estimators_):
            if est.n_trees_per_iteration_ == 1:
                names.extend(f"gb{i}_tree{k}" for k in range(est.n_estimators))
            else:
                for j in range(est.n_trees_per_iteration_):
                    names.extend(f"gb{i}_cls{j}_tree{k}" for k in range(est.n_estimators))
        return np.asarray(names, dtype=object)
