"""Same import surface as the reference's ``ais_benchmarks.distributions`` for the classes on the hot path
(reference ais_benchmarks/distributions/__init__.py:1-7)."""
from .distributions import CDistribution
from .parametric import CMultivariateNormal, CMultivariateUniform
from .mixture import CMixtureModel, CGaussianMixtureModel

__all__ = ["CDistribution", "CMultivariateNormal", "CMultivariateUniform", "CMixtureModel", "CGaussianMixtureModel"]
