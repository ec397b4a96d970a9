"""Multivariate Normal with the reference's constructor and attributes
(reference ais_benchmarks/distributions/parametric/CMultivariateNormal.py:26-74); the density is
evaluated by the sm_100a mixture kernel with K = 1 (C ABI aisb_mixture_logpdf_f32).

"sigma" is the COVARIANCE matrix, not a standard deviation.
"""
import numpy as np
import torch

from ... import _device
from ..._lib import COV_DIAG, COV_FULL
from ..distributions import CDistribution


class CMultivariateNormal(CDistribution):
    def __init__(self, params):
        self._check_param(params, "mean")
        self._check_param(params, "sigma")

        params["type"] = "normal"
        params["family"] = "exponential"
        params["likelihood_f"] = self.prob
        params["loglikelihood_f"] = self.log_prob
        params["dims"] = len(params["mean"])

        # 6-"sigma" clipped support, computed exactly like CMultivariateNormal.py:37-40
        if "support" not in params:
            mean = np.asarray(params["mean"], dtype=np.float64)
            sd = np.sqrt(np.diag(np.asarray(params["sigma"], dtype=np.float64)))
            params["support"] = np.array([mean - sd * 6.0, mean + sd * 6.0])

        super(CMultivariateNormal, self).__init__(params)
        self._version = 0
        self._packed = None
        self.set_moments(np.array(params["mean"], dtype=np.float64), np.array(params["sigma"], dtype=np.float64))
        assert self._scale.shape == (self.dims, self.dims)
        self.term1 = - 0.5 * self.dims * np.log(np.pi * 2)

    def set_moments(self, mean, cov):
        """CMultivariateNormal.py:52-62: asserts the shape and det > 0, precomputes inv / log det."""
        mean = np.asarray(mean, dtype=np.float64)
        cov = np.asarray(cov, dtype=np.float64)
        assert cov.shape == (self.dims, self.dims)
        self.det = np.linalg.det(cov)
        assert self.det > 0

        self.loc = mean
        self.scale = cov
        self.inv_cov = np.linalg.inv(cov)
        self.log_det = np.log(self.det)
        self.term2 = - 0.5 * self.log_det
        self._version += 1
        self._packed = None

    # -- parameter views used by the mixture classes to pack many components at once -----------
    def is_diagonal(self):
        c = self._scale
        return bool(np.count_nonzero(c - np.diag(np.diagonal(c))) == 0)

    def _device_mixture(self):
        if self._packed is None:
            mean = self._loc.reshape(1, self.dims)
            if self.is_diagonal():
                self._packed = _device.DeviceMixture.from_params(mean, np.diagonal(self._scale).reshape(1, -1), None,
                                                                 COV_DIAG)
            else:
                self._packed = _device.DeviceMixture.from_params(mean, self._scale.reshape(1, self.dims, self.dims),
                                                                 None, COV_FULL)
        return self._packed

    def sample(self, n_samples=1, as_tensor=False):
        """n_samples draws, shape (n, dims) (CMultivariateNormal.py:64-65). Drawn on the device with the
        Philox kernel (RNG streams are not comparable with np.random.multivariate_normal)."""
        _device.require_cuda()
        dev = _device.device()
        z = _device.sample_iso(torch.zeros((1, self.dims), dtype=torch.float32, device=dev), int(n_samples), 1.0)
        if self.is_diagonal():
            sd = torch.from_numpy(np.sqrt(np.diagonal(self._scale))).to(dev, torch.float32)
            x = z * sd
        else:
            L = torch.from_numpy(np.linalg.cholesky(0.5 * (self._scale + self._scale.T))).to(dev, torch.float32)
            x = z @ L.T
        x = x + torch.from_numpy(self._loc.reshape(1, -1)).to(dev, torch.float32)
        return x if as_tensor else _device.download(x)

    def log_prob(self, samples):
        x, as_numpy = _device.to_device_points(samples, self.dims)
        return _device.from_device(_device.mixture_logpdf(x, self._device_mixture()), as_numpy)

    def prob(self, samples):
        # exp(log_prob), CMultivariateNormal.py:73-74; exponentiated on the device in float64 for numpy callers
        x, as_numpy = _device.to_device_points(samples, self.dims)
        return _device.exp_from_device(_device.mixture_logpdf(x, self._device_mixture()), as_numpy)

    def condition(self, dist):
        raise NotImplementedError

    def marginal(self, dim):
        raise NotImplementedError
