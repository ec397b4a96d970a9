from .base import CMetric
from .divergences import CDivergence, CKLDivergence

__all__ = ["CMetric", "CDivergence", "CKLDivergence"]
