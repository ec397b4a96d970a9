from .base import CMetric
from .divergences import CDivergence, CJSDivergence, CKLDivergence
from .statistics import CExpectedValueMSE

__all__ = ["CMetric", "CDivergence", "CKLDivergence", "CJSDivergence", "CExpectedValueMSE"]
