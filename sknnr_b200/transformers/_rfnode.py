"""Random-forest node transformer for RFNN (mirrors
ref:src/sknnr/transformers/_rfnode_transformer.py); fit / transform live in ``_treenode``."""

from __future__ import annotations

import numpy as np
from sklearn.ensemble import RandomForestClassifier, RandomForestRegressor
from sklearn.utils.validation import check_is_fitted

from ._treenode import TreeNodeTransformer


class RFNodeTransformer(TreeNodeTransformer):
    def __init__(self, n_estimators=50, criterion_reg="squared_error", criterion_clf="gini",
                 max_depth=None, min_samples_split=2, min_samples_leaf=5,
                 min_weight_fraction_leaf=0.0, max_features_reg=1.0, max_features_clf="sqrt",
                 max_leaf_nodes=None, min_impurity_decrease=0.0, bootstrap=True, oob_score=False,
                 n_jobs=None, random_state=None, verbose=0, warm_start=False,
                 class_weight_clf=None, ccp_alpha=0.0, max_samples=None, monotonic_cst=None):
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

    def fit(self, X, y):
        common = dict(
            n_estimators=self.n_estimators, max_depth=self.max_depth,
            min_samples_split=self.min_samples_split, min_samples_leaf=self.min_samples_leaf,
            min_weight_fraction_leaf=self.min_weight_fraction_leaf,
            max_leaf_nodes=self.max_leaf_nodes, min_impurity_decrease=self.min_impurity_decrease,
            bootstrap=self.bootstrap, oob_score=self.oob_score, n_jobs=self.n_jobs,
            random_state=self.random_state, verbose=self.verbose, warm_start=self.warm_start,
            ccp_alpha=self.ccp_alpha, max_samples=self.max_samples, monotonic_cst=self.monotonic_cst)
        return self._fit(
            X, y,
            lambda: RandomForestRegressor(criterion=self.criterion_reg, max_features=self.max_features_reg, **common),
            lambda: RandomForestClassifier(criterion=self.criterion_clf, max_features=self.max_features_clf,
                                           class_weight=self.class_weight_clf, **common))

    def _set_n_trees_per_iteration(self):
        return [1] * self.n_forests_

    def _set_tree_weights(self, X, y):
        return [np.full(self.n_estimators, 1.0 / self.n_estimators, dtype=np.float64)
                for _ in range(self.n_forests_)]

    def _trees(self):
        return [t.tree_ for est in self.estimators_ for t in est.estimators_]

    def get_feature_names_out(self, input_features=None):
        check_is_fitted(self, "estimators_")
        return np.asarray([f"rf{i}_tree{j}" for i, e in enumerate(self.estimators_)
                           for j in range(e.n_estimators)], dtype=object)
