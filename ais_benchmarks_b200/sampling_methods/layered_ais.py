"""Layered AIS with the reference's interface
(reference ais_benchmarks/sampling_methods/layered_ais.py:10-107).

Lower layer: identical DM weighting to DM-AIS (one fused kernel per iteration). Upper layer: the N
proposal means are N independent Metropolis-Hastings chains advanced L steps per iteration by ONE
kernel launch (C ABI aisb_lais_mh_f32) instead of N*L*2 scalar target evaluations.
"""
import time

import numpy as np
import torch

from .. import _device
from .base import target_device_mixture
from .dm_ais import CDeterministicMixtureAIS


class CLayeredAIS(CDeterministicMixtureAIS):
    def __init__(self, params):
        """
        :param params: as CDeterministicMixtureAIS plus
            - L: MH steps per adaptation; mh_sigma: VARIANCE of the MH random-walk proposal (layered_ais.py:31-32)
        """
        self.L = params["L"]
        self.mhsigma = params["mh_sigma"]
        super(CLayeredAIS, self).__init__(params)

    def mh_noise(self):
        """(delta [L, N, D] ~ N(0, mh_sigma I), u [L, N] ~ U(0,1)) on the device."""
        dev = _device.device()
        zeros = torch.zeros((1, self.ndims), dtype=torch.float32, device=dev)
        delta = _device.sample_iso(zeros, int(self.L) * int(self.N), float(self.mhsigma))
        return delta.reshape(self.L, self.N, self.ndims), torch.rand((self.L, self.N), dtype=torch.float32, device=dev)

    def mcmc_step(self, target_d, delta, u):
        """Advance every chain L steps with caller-supplied noise (layered_ais.py:54-74)."""
        target = target_device_mixture(target_d)
        if target is None:
            raise NotImplementedError("CLayeredAIS on B200 needs a Gaussian / GMM target from this package for the "
                                      "on-device MH layer")
        dev = _device.device()
        means = self.proposal_means().clone()
        lo = torch.from_numpy(self.space_min).to(dev, torch.float32)
        hi = torch.from_numpy(self.space_max).to(dev, torch.float32)
        _device.lais_mh(means, target, delta.contiguous(), u.contiguous(), lo, hi)
        self._num_pi_evals += 2 * self.L * self.N
        self._num_q_samples += self.L * self.N
        self._set_means_dev(means)

    def resample(self, target_d):
        delta, u = self.mh_noise()
        self.mcmc_step(target_d, delta, u)

    def importance_sample(self, target_d, n_samples, timeout=60):
        elapsed_time = 0
        t_ini = time.time()
        while n_samples > self._n() and elapsed_time < timeout:
            new_samples = _device.sample_iso(self.proposal_means(), int(self.K), float(self.sigma))
            self._num_q_samples += self.N * self.K
            new_logw, _ = self.weigh(new_samples, target_d, hint=self._own_proposal())
            self._num_pi_evals += self.N * self.K
            self.resample(target_d)
            self._append(new_samples, new_logw)
            elapsed_time = time.time() - t_ini
        norm_w = _device.normalize_weights(self._logw_dev, self._weight_partials())
        return self._out(self._samples_dev), (norm_w if self.device_outputs else norm_w.double().cpu().numpy())
