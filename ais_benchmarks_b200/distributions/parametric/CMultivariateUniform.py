"""Box-uniform distribution with the reference's constructor
(reference ais_benchmarks/distributions/parametric/CMultivariateUniform.py:5-59). Density is the
sm_100a box-mixture kernel with one box (C ABI aisb_box_mixture_logpdf_f32); the inlier test is
the reference's  min < x <= max  (quirk Q6).
"""
import numpy as np
import torch

from ... import _device
from ..distributions import CDistribution


class CMultivariateUniform(CDistribution):
    def __init__(self, params):
        self._check_param(params, "center")
        self._check_param(params, "radius")

        params["type"] = "uniform"
        params["family"] = "uniform"
        params["likelihood_f"] = self.prob
        params["loglikelihood_f"] = self.log_prob
        params["dims"] = len(params["center"])
        params["center"] = np.asarray(params["center"], dtype=np.float64)
        params["radius"] = np.asarray(params["radius"], dtype=np.float64)
        params["support"] = [params["center"] - params["radius"], params["center"] + params["radius"]]
        super(CMultivariateUniform, self).__init__(params)
        self.loc = params["center"]
        self.scale = params["radius"]
        self.dims = len(self._loc)
        # one radius for all dimensions, or one per dimension (CMultivariateUniform.py:20-27)
        if len(self._scale) == 1:
            self.volume = (self._scale * 2) ** self.dims
        elif len(self._scale) == self.dims:
            self.volume = np.prod([self._scale * 2])
        else:
            raise ValueError("Radius dimensionality not valid. It has to be 1 dim (all dimensions will use the same "
                             "radius) or N dim, where N is the dimensionality")
        self.prob_val = 1 / self.volume
        self.logprob_val = np.log(self.prob_val)
        self._box = None

    def box(self):
        """(lo [D], hi [D], log(1/volume)) as float64 host values."""
        c = np.asarray(self._loc, dtype=np.float64).reshape(-1)
        r = np.broadcast_to(np.asarray(self._scale, dtype=np.float64).reshape(-1), (self.dims,))
        lo, hi = c - r, c + r
        return lo, hi, float(np.asarray(self.logprob_val).reshape(-1)[0])

    def _device_box(self):
        if self._box is None:
            _device.require_cuda()
            lo, hi, lpv = self.box()
            dev = _device.device()
            self._box = (torch.from_numpy(lo.reshape(1, -1)).to(dev, torch.float32),
                         torch.from_numpy(hi.reshape(1, -1)).to(dev, torch.float32),
                         torch.tensor([lpv], dtype=torch.float32, device=dev))
        return self._box

    def sample(self, n_samples=1, as_tensor=False):
        lo, hi, _ = self._device_box()
        x = _device.uniform_box(lo[0].contiguous(), hi[0].contiguous(), int(n_samples))
        return x if as_tensor else _device.download(x)

    def log_prob(self, samples):
        x, as_numpy = _device.to_device_points(samples, self.dims)
        lo, hi, lwv = self._device_box()
        return _device.from_device(_device.box_mixture_logpdf(x, lo, hi, lwv, open_upper=False), as_numpy)

    def prob(self, samples):
        x, as_numpy = _device.to_device_points(samples, self.dims)
        lo, hi, lwv = self._device_box()
        return _device.exp_from_device(_device.box_mixture_logpdf(x, lo, hi, lwv, open_upper=False), as_numpy)

    def condition(self, dist):
        raise NotImplementedError

    def integral(self, a, b, nsamples=None):
        assert np.all(a < b)
        ini = np.maximum(a, self.support_vals[0])
        end = np.minimum(b, self.support_vals[1])
        return (1 / self.volume) * np.prod((end - ini))

    def marginal(self, dim):
        raise NotImplementedError
