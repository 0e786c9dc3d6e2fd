from ._base import ComponentReducerMixin, StandardScalerWithDOF
from ._float_transformers import CCATransformer, CCorATransformer, MahalanobisTransformer
from ._gbnode import GBNodeTransformer
from ._rfnode import RFNodeTransformer
from ._treenode import TreeNodeTransformer

__all__ = [
    "StandardScalerWithDOF", "MahalanobisTransformer", "CCATransformer", "CCorATransformer",
    "RFNodeTransformer", "GBNodeTransformer", "TreeNodeTransformer", "ComponentReducerMixin",
]
