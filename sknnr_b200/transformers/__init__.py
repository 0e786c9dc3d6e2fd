from ._base import ComponentReducerMixin, StandardScalerWithDOF
from ._float_transformers import CCATransformer, CCorATransformer, MahalanobisTransformer
from ._rfnode import RFNodeTransformer

__all__ = [
    "StandardScalerWithDOF", "MahalanobisTransformer", "CCATransformer", "CCorATransformer",
    "RFNodeTransformer", "ComponentReducerMixin",
]
