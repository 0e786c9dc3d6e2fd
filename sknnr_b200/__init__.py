"""sknnr_b200: B200-native query-time hot path of lemma-osu/sknnr behind the unchanged
estimator surface (Raw / Euclidean / Mahalanobis / MSN / GNN / RFNN / GBNN regressors).

The compute path is hand-written CUDA for sm_100a in ``csrc/`` behind the C ABI of
``include/sknnr_b200.h``; importing the package does not need a GPU, running a query does.
"""

from ._base import RawKNNRegressor
from ._estimators import (
    EuclideanKNNRegressor,
    GBNNRegressor,
    GNNRegressor,
    MahalanobisKNNRegressor,
    MSNRegressor,
    RFNNRegressor,
)

from .raster import kneighbors_raster, predict_raster

__version__ = "0.1.0"

__all__ = [
    "RawKNNRegressor", "EuclideanKNNRegressor", "MahalanobisKNNRegressor", "MSNRegressor",
    "GNNRegressor", "RFNNRegressor", "GBNNRegressor", "kneighbors_raster", "predict_raster",
]
