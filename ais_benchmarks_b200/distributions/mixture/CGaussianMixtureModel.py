"""Diagonal-covariance Gaussian mixture with the reference's constructor
(reference ais_benchmarks/distributions/mixture/CGaussianMixtureModel.py:8-63).

"sigmas" holds per-dimension VARIANCES that go on the covariance diagonal (CGaussianMixtureModel.py:30);
``self.weights`` keeps the raw user weights, the normalised copy lives in ``self.gmm.weights``.
"""
import numpy as np

from ..distributions import CDistribution
from ..parametric.CMultivariateNormal import CMultivariateNormal
from .CMixtureModel import CMixtureModel


class CGaussianMixtureModel(CDistribution):
    def __init__(self, params):
        self._check_param(params, "means")
        self._check_param(params, "sigmas")

        params["type"] = "GMM"
        params["family"] = "mixture"
        params["likelihood_f"] = self.prob
        params["loglikelihood_f"] = self.log_prob
        params["dims"] = len(params["means"][0])
        super(CGaussianMixtureModel, self).__init__(params)

        self.models = []
        self.means = params["means"]
        self.sigmas = params["sigmas"]
        for mean, cov in zip(self.means, self.sigmas):
            self.models.append(CMultivariateNormal({"mean": np.asarray(mean, dtype=np.float64),
                                                    "sigma": np.diag(np.asarray(cov, dtype=np.float64))}))

        if "weights" in params.keys():
            self.weights = params["weights"]
        else:
            self.weights = (1.0 / len(self.means)) * np.array([1] * len(self.means))

        # the reference hands the inner mixture the bound method self.support (CGaussianMixtureModel.py:38, never read);
        # the support values are passed instead so that the object is not a reference cycle
        self.gmm = CMixtureModel({"models": self.models, "weights": self.weights,
                                  "dims": self.dims, "support": self.support_vals})

    def log_prob(self, samples):
        return self.gmm.log_prob(samples)

    def prob(self, samples):
        return self.gmm.prob(samples)

    def sample(self, nsamples=1, **kwargs):
        return self.gmm.sample(nsamples, **kwargs)

    def device_mixture(self):
        """The packed HBM-resident parameters (used by the fused DM-weight kernel)."""
        kind, packed = self.gmm._packed()
        return packed

    def condition(self, dist):
        raise NotImplementedError

    def marginal(self, dim):
        raise NotImplementedError

    def integral(self, a, b, nsamples=None):
        raise NotImplementedError

    def to_dict(self, name=None, batch_size=32, nsamples=1000, nsamples_eval=2000):
        res = super(CGaussianMixtureModel, self).to_dict(name, batch_size, nsamples, nsamples_eval)
        res["params"]["means"] = [np.asarray(v).tolist() for v in self.means]
        res["params"]["sigmas"] = [np.asarray(v).tolist() for v in self.sigmas]
        res["params"]["weights"] = np.asarray(self.weights).tolist()
        return res
