from .base import CMetric
from .divergences import CDivergence, CJSDivergence, CKLDivergence
from .performance import CElapsedTime, CMemoryUsage
from .statistics import CExpectedValueMSE

__all__ = ["CMetric", "CDivergence", "CKLDivergence", "CJSDivergence", "CExpectedValueMSE", "CElapsedTime",
           "CMemoryUsage"]
