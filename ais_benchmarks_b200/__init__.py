"""ais_benchmarks_b200 -- B200 (sm_100a) implementation of the ais-benchmarks density /
importance-weight / KLD hot path, behind the reference's batch-first Python API.

    from ais_benchmarks_b200.distributions import CGaussianMixtureModel, CMultivariateNormal, CMixtureModel
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CMixturePMC, CLayeredAIS
    from ais_benchmarks_b200.metrics import CKLDivergence

Arithmetic runs in hand-written CUDA kernels reached through the C ABI of libais_b200.so
(include/ais_b200.h); there is no CPU fallback.
"""
from . import _lib
from ._lib import AisbError

__version__ = "0.1.0"


def library_path():
    return _lib.LIB_PATH
