"""Same import surface as the reference's ``ais_benchmarks.sampling_methods`` for the samplers on the hot path
(reference ais_benchmarks/sampling_methods/__init__.py:1-9)."""
from .base import CSamplingMethod, CMixtureSamplingMethod, CMixtureISSamplingMethod, CDeviceKDE
from .dm_ais import CDeterministicMixtureAIS
from .layered_ais import CLayeredAIS
from .m_pmc import CMixturePMC
from .tree_pyramid import CTreePyramid, CTreePyramidSampling

__all__ = ["CSamplingMethod", "CMixtureSamplingMethod", "CMixtureISSamplingMethod", "CDeviceKDE",
           "CDeterministicMixtureAIS", "CLayeredAIS", "CMixturePMC", "CTreePyramid", "CTreePyramidSampling"]
