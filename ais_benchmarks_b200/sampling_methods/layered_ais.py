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

    def mh_noise(self, n_chains=None):
        """(delta [L, n, D] ~ N(0, mh_sigma I), u [L, n] ~ U(0,1)) on the device for n chains (default: all N)."""
        dev = _device.device()
        n = int(self.N if n_chains is None else n_chains)
        zeros = torch.zeros((1, self.ndims), dtype=torch.float32, device=dev)
        delta = _device.sample_iso(zeros, int(self.L) * n, float(self.mhsigma))
        return delta.reshape(self.L, n, self.ndims), torch.rand((self.L, n), dtype=torch.float32, device=dev)

    def mcmc_step(self, target_d, delta, u, chains=None):
        """Advance the chains [chains[0], chains[1]) (default: all) L steps with caller-supplied noise
        (layered_ais.py:54-74). Sharded: every rank advances its own slice of the N independent chains, then the N x D
        means are all-gathered."""
        target = target_device_mixture(target_d)
        if target is None:
            raise NotImplementedError("CLayeredAIS on B200 needs a Gaussian / GMM target from this package for the "
                                      "on-device MH layer")
        dev = _device.device()
        a, b = (0, int(self.N)) if chains is None else chains
        means = self.proposal_means()[a:b].clone()
        lo = torch.from_numpy(self.space_min).to(dev, torch.float32)
        hi = torch.from_numpy(self.space_max).to(dev, torch.float32)
        if b > a:
            _device.lais_mh(means, target, delta.contiguous(), u.contiguous(), lo, hi)
        self._num_pi_evals += 2 * self.L * (b - a)
        self._num_q_samples += self.L * (b - a)
        if chains is not None:
            from .. import parallel
            _, ws = parallel.world()
            counts = [e - s for s, e in (parallel.shard_rows(int(self.N), r, ws) for r in range(ws))]
            means = parallel.all_gather_rows(means, counts)
        self._set_means_dev(means)

    def _prepare_graph(self, dev):
        self._graph_consts = (torch.zeros((1, self.ndims), dtype=torch.float32, device=dev),
                              torch.from_numpy(self.space_min.astype(np.float32)).to(dev),
                              torch.from_numpy(self.space_max.astype(np.float32)).to(dev))

    def _adapt_in_graph(self, mbuf, x, logw, target):
        """Captured-iteration counterpart of resample(): L MH steps of all N chains, in place on the means."""
        zeros, lo, hi = self._graph_consts
        delta = _device.sample_iso_ctr(zeros, int(self.L) * int(self.N), float(self.mhsigma))
        u = torch.rand((int(self.L), int(self.N)), dtype=torch.float32, device=mbuf.device)
        _device.lais_mh(mbuf, target, delta.reshape(self.L, self.N, self.ndims), u, lo, hi)
        self._graph_mh_keep = (delta, u)

    def _graph_iteration(self, target_d, per):
        if not super(CLayeredAIS, self)._graph_iteration(target_d, per):
            return False
        self._num_pi_evals += 2 * self.L * self.N
        self._num_q_samples += self.L * self.N
        return True

    def resample(self, target_d):
        if self.sharded:
            from .. import parallel
            a, b = parallel.shard_rows(int(self.N))
            delta, u = self.mh_noise(b - a)
            self.mcmc_step(target_d, delta, u, chains=(a, b))
            return
        delta, u = self.mh_noise()
        self.mcmc_step(target_d, delta, u)

    def importance_sample(self, target_d, n_samples, timeout=60):
        elapsed_time = 0
        t_ini = time.time()
        per = self._draws_per_proposal()
        # (sharded: the loop is driven by the sample count alone -- a per-rank clock must not desynchronise the collectives)
        while n_samples > (self._n_all_ranks if self.sharded else self._n()) and (self.sharded or elapsed_time < timeout):
            if self._graph_wanted(target_d, per) and self._graph_iteration(target_d, per):
                elapsed_time = time.time() - t_ini
                continue
            new_samples = _device.sample_iso(self.proposal_means(), per, float(self.sigma))
            self._num_q_samples += self.N * per
            new_logw, _ = self.weigh(new_samples, target_d, hint=self._own_proposal(per))
            self._num_pi_evals += self.N * per
            self.resample(target_d)
            self._append(new_samples, new_logw)
            self._n_all_ranks += self.N * int(self.K)
            elapsed_time = time.time() - t_ini
        self._check_pending_records()
        norm_w = _device.normalize_weights(self._logw_dev, self._global_partials())
        return self._out(self._samples_dev), (norm_w if self.device_outputs else _device.download(norm_w))
