import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
from ais_benchmarks_b200.metrics import CKLDivergence
target = CGaussianMixtureModel({"means": [[-0.5, -0.4], [0.5, 0.3]], "sigmas": [[0.02, 0.02], [0.03, 0.01]],
                                "weights": [0.4, 0.6], "support": [[-1, -1], [1, 1]]})
method = CDeterministicMixtureAIS({"space_min": np.array([-1., -1.]), "space_max": np.array([1., 1.]), "dims": 2,
                                   "K": 10, "N": 5, "J": 1000, "sigma": 0.05})
samples, weights = method.importance_sample(target, n_samples=1000)
kld = CKLDivergence().compute(p=target, q=method, nsamples=2000)
print("snippet ok", samples.shape, weights.shape, float(weights.sum()), kld, method.get_NESS(), method.get_acceptance_rate())
