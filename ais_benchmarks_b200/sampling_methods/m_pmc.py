"""Mixture PMC with the reference's interface
(reference ais_benchmarks/sampling_methods/m_pmc.py:10-121).

Per iteration: K draws from the proposal mixture, one fused weight kernel (log pi, log q, log w,
normaliser), one Rao-Blackwellised moment kernel (alpha_d and mu_d numerators over all K x N
(sample, proposal) pairs; C ABI aisb_mpmc_moments[_dev]_f32), one raw-moment kernel for the scatter and ONE
update kernel (C ABI aisb_mpmc_update_f64: alpha / mu / Sigma, failsafe, determinant check, weight sanitising
and the re-packing of the proposal mixture, all in float64 on the device). The proposal parameters live in
HBM between iterations; the host reads 16 bytes of status per iteration (which packed layout to use, did a
covariance lose positive definiteness) and mirrors the parameters as numpy arrays only when somebody asks
(`proposals`, `model`, `wproposals`, ...). D > 16 and `paper_covariance=True` keep the float64 host update.

Reference quirks kept on purpose: Q3 - the covariance update multiplies the UNWEIGHTED scatter
(X - mu_d)^T (X - mu_d) by sum_i w_i rho_di / alpha_d = 1, and falls back to 10*I when any entry is NaN or > 10
(m_pmc.py:64-72); Q4 - returned weights are normalised per batch, not globally (m_pmc.py:108,118).
``paper_covariance=True`` in params switches to the weighted covariance of eq. 14.3 instead (not parity mode; D <= 8:
the weighted second moments come from the same component-stationary kernel, C ABI aisb_mpmc_moments2_f32).
"""
import time

import numpy as np
import torch

from .. import _device
from .._lib import COV_FULL
from ..distributions import CMixtureModel, CMultivariateNormal
from ..distributions.mixture.CMixtureModel import sanitize_weights
from .base import (CMixtureISSamplingMethod, foreign_target_logp, mixture_component_draws,
                   target_device_mixture)


class CMixturePMC(CMixtureISSamplingMethod):
    def __init__(self, params):
        super(CMixturePMC, self).__init__(params)
        self.K = params["K"]            # samples per iteration (paper: N)
        self.N = params["N"]            # mixture components (paper: D)
        self.J = params["J"]
        self.sigma = params["sigma"]
        self.paper_covariance = bool(params.get("paper_covariance", False))
        if self.paper_covariance and params["dims"] > 8:
            raise NotImplementedError("paper_covariance is limited to dims <= 8 (second-moment kernel)")
        # sharded (under torch.distributed): every rank draws and weighs its share of the K samples of an iteration;
        # the sufficient statistics (weight record, alpha / mu numerators, raw moments) are summed over the ranks and
        # every rank performs the identical float64 parameter update
        self.sharded = bool(params.get("sharded", False))
        # False keeps the float64 host update of round 1 (also taken for D > 16 and paper_covariance)
        self.device_update = bool(params.get("device_update", True))
        self._dev = None
        self._host = {}
        self.reset()

    def reset(self):
        super(CMixturePMC, self).reset()
        means = np.random.uniform(self.space_min, self.space_max, size=(self.N, self.ndims))
        if self.sharded:
            from .. import parallel
            means = parallel.broadcast_array(means)
        self._n_all_ranks = 0
        self._dev_valid = False
        self._host = {"means": means,
                      "covs": np.tile(np.diag(np.array([self.sigma] * self.ndims, dtype=np.float64)), (self.N, 1, 1)),
                      "alpha": np.array([1 / self.N] * self.N)}
        self._host["wmix"] = sanitize_weights(self._host["alpha"].copy())
        self._mixture = None
        self._objects = None
        self._sampler_params = None

    # -- proposal parameters: numpy float64 mirrors of the device state, pulled on demand ----------------
    def _pull(self):
        if self._dev_valid and self._host is None:
            d = self._dev
            self._host = {"means": d["means"].cpu().numpy(), "covs": d["covs"].cpu().numpy(),
                          "alpha": d["alpha"].cpu().numpy(), "wmix": d["wmix"].cpu().numpy()}
        return self._host

    def _set_host(self, key, val):
        self._pull()[key] = val
        self._dev_valid = False            # the host copy is the truth again: re-pack from it
        self._mixture = None
        self._objects = None
        self._sampler_params = None

    _means = property(lambda self: self._pull()["means"], lambda self, v: self._set_host("means", v))
    _covs = property(lambda self: self._pull()["covs"], lambda self, v: self._set_host("covs", v))
    wproposals = property(lambda self: self._pull()["alpha"], lambda self, v: self._set_host("alpha", v))
    _mix_weights = property(lambda self: self._pull()["wmix"], lambda self, v: self._set_host("wmix", v))

    # -- proposal mixture --------------------------------------------------------------------------
    def proposal_mixture(self):
        if self._mixture is None and self._dev_valid:
            d = self._dev
            from .._lib import COV_DIAG
            kind = COV_DIAG if d["all_diag"] else COV_FULL
            self._mixture = _device.DeviceMixture(d["rows_diag"] if d["all_diag"] else d["rows_full"], d["offsets"],
                                                  self.N, self.ndims, kind)
        if self._mixture is None:
            diag = np.diagonal(self._covs, axis1=1, axis2=2)
            if np.count_nonzero(self._covs - diag[:, :, None] * np.eye(self.ndims)[None]) == 0:
                from .._lib import COV_DIAG
                self._mixture = _device.DeviceMixture.from_params(self._means, np.ascontiguousarray(diag),
                                                                  self._mix_weights, COV_DIAG)
            else:
                self._mixture = _device.DeviceMixture.from_params(self._means, self._covs, self._mix_weights, COV_FULL)
        return self._mixture

    def _build_objects(self):
        if self._objects is None:
            props = [CMultivariateNormal({"mean": m.copy(), "sigma": c.copy()}) for m, c in zip(self._means, self._covs)]
            model = CMixtureModel({"models": props, "weights": self.wproposals.copy(),
                                   "dims": self.ndims, "support": [self.space_min, self.space_max]})
            self._objects = (props, model)
        return self._objects

    @property
    def proposals(self):
        return self._build_objects()[0]

    @proposals.setter
    def proposals(self, val):
        pass

    @property
    def model(self):
        return self._build_objects()[1]

    @model.setter
    def model(self, val):
        pass

    def log_prob(self, s):
        x, as_numpy = _device.to_device_points(s, self.ndims)
        return _device.from_device(_device.mixture_logpdf(x, self.proposal_mixture()), as_numpy)

    def prob(self, s):
        x, as_numpy = _device.to_device_points(s, self.ndims)
        return _device.exp_from_device(_device.mixture_logpdf(x, self.proposal_mixture()), as_numpy)

    def _sample_device(self, n):
        """Ancestral sampling from the proposal mixture on the device (m_pmc.py:86 -> CMixtureModel.sample): component
        indices from a device multinomial, then ONE kernel for x = mu_k + L_k z (C ABI aisb_sample_mixture_f32). The
        Cholesky factors are computed on the host in float64 once per parameter update (N small D x D matrices)."""
        dev = _device.device()
        if self._dev_valid:
            d = self._dev
            if d["all_zero"]:    # degenerate adaptation: the reference draws every sample from the LAST component
                idx = torch.full((int(n),), self.N - 1, dtype=torch.int64, device=dev)
            else:
                idx = torch.multinomial(d["w32"], int(n), replacement=True)
            return _device.sample_mixture(d["means32"], d["chol32"], idx.contiguous())
        if self._sampler_params is None:
            L = np.linalg.cholesky(0.5 * (self._covs + np.transpose(self._covs, (0, 2, 1))))
            self._sampler_params = (torch.from_numpy(self._means).to(dev, torch.float32).contiguous(),
                                    torch.from_numpy(L).to(dev, torch.float32).contiguous())
        idx = mixture_component_draws(self._mix_weights, n, dev)
        return _device.sample_mixture(self._sampler_params[0], self._sampler_params[1], idx.contiguous())

    def sample(self, n_samples):
        self._num_q_samples += n_samples
        x = self._sample_device(n_samples)
        return x if self.device_outputs else _device.download(x)

    # -- one M-PMC iteration on a given batch (also the parity-test entry point) -------------------
    def weigh(self, new_samples, target_d):
        """-> (log w [K], log q [K], partials) for a batch (m_pmc.py:93-96)."""
        target = target_device_mixture(target_d)
        logp_in = None if target is not None else foreign_target_logp(target_d, new_samples)
        logw, _, logq, partials = _device.dm_weights(new_samples, target, logp_in, self.proposal_mixture(),
                                                    want_logq=True)
        return logw, logq, partials

    # -- the update on the device ------------------------------------------------------------------------
    def _device_update_ok(self):
        return self.device_update and not self.paper_covariance and self.ndims <= 16

    def _device_buffers(self):
        if self._dev is None:
            lib = _device.require_cuda()
            dev, N, D = _device.device(), int(self.N), int(self.ndims)
            KP, DP = lib.aisb_padded_components(N), lib.aisb_padded_dims(D)
            RF = lib.aisb_packed_row_floats(COV_FULL, D)
            f64 = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
            f32 = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
            self._dev = {"mom": f64(N, D + 1), "raw": f64(D + D * D), "means": f64(N, D), "covs": f64(N, D, D),
                         "alpha": f64(N), "wmix": f64(N), "rows_diag": f32(KP, 2 * DP), "rows_full": f32(KP, RF),
                         "offsets": f32(KP), "means32": f32(N, D), "chol32": f32(N, D, D), "w32": f32(N),
                         "status": torch.zeros(4, dtype=torch.int32, device=dev), "all_diag": True, "all_zero": False}
        return self._dev

    def adapt_device(self, samples, logw, logq, record, n_total):
        """m_pmc.py:53-78 without leaving HBM: moment kernels -> (all-reduce over the ranks when sharded) -> one update
        kernel that also re-packs the proposal mixture. `record`: float64 CUDA [4] weight record of the whole batch
        (merged over the ranks when sharded). One 16-byte status read-back per call."""
        lib = _device.require_cuda()
        d = self._device_buffers()
        mix = self.proposal_mixture()                     # the mixture the samples were drawn from / weighed against
        d["mom"].zero_()
        d["raw"].zero_()
        p = _device._ptr
        _device.check(lib.aisb_mpmc_moments_dev_f32(p(samples), samples.shape[0], mix.ref(), p(logq), p(logw),
                                                    p(record), p(d["mom"]), _device.stream_ptr()),
                      "aisb_mpmc_moments_dev_f32")
        _device.check(lib.aisb_raw_moments_f32(p(samples), samples.shape[0], self.ndims, p(d["raw"]),
                                               _device.stream_ptr()), "aisb_raw_moments_f32")
        if self.sharded:
            from .. import parallel
            parallel.all_reduce_sum(d["mom"])
            parallel.all_reduce_sum(d["raw"])
        # the moment kernel above was the last reader of the old packed rows: the update may overwrite them in place
        _device.check(lib.aisb_mpmc_update_f64(p(d["mom"]), p(d["raw"]), float(n_total), int(self.N), int(self.ndims),
                                               p(d["means"]), p(d["covs"]), p(d["alpha"]), p(d["wmix"]),
                                               p(d["rows_diag"]), p(d["rows_full"]), p(d["offsets"]), p(d["means32"]),
                                               p(d["chol32"]), p(d["w32"]), p(d["status"]), _device.stream_ptr()),
                      "aisb_mpmc_update_f64")
        st = d["status"].cpu().numpy()                    # 16 bytes: the one host read of the iteration
        # CMultivariateNormal.set_moments asserts det > 0 (CMultivariateNormal.py:55-56)
        assert st[2] == 0, "%d proposal covariance(s) lost positive definiteness in the M-PMC update" % int(st[2])
        d["all_diag"], d["all_zero"] = bool(st[1] == 0), bool(st[3] != 0)
        self._dev_valid = True
        self._host = None                                  # numpy mirrors are pulled when somebody reads them
        self._mixture = None
        self._objects = None
        self._sampler_params = None

    def adapt(self, samples, logw, logq, partials_host):
        """alpha, mu, Sigma update (m_pmc.py:53-78) from device-side sufficient statistics, finished on the HOST in
        float64 (the round-1 path; still used for D > 16, paper_covariance and as the parity entry point of the tests)."""
        log_norm = _device.weight_logsum(partials_host)  # log sum_i w_i: weights are batch-normalised (m_pmc.py:108)
        D = self.ndims
        mom = _device.mpmc_moments(samples, self.proposal_mixture(), logq, logw, log_norm,
                                   second=self.paper_covariance).cpu().numpy()
        if self.sharded:
            from .. import parallel
            mom = parallel.all_reduce_sum_host(mom)
        alpha = mom[:, 0].copy()                           # eq. 14.1 (Rao-Blackwellised)
        with np.errstate(divide="ignore", invalid="ignore"):
            mu = mom[:, 1:D + 1] / alpha[:, None]          # eq. 14.2
        if self.paper_covariance:
            self._finish_adapt(alpha, mu, self._paper_covariances(mom, alpha, mu))
            return
        n = samples.shape[0]
        s1_t, s2_t = _device.raw_moments(samples)
        s1, s2 = s1_t.cpu().numpy(), s2_t.cpu().numpy()
        if self.sharded:
            flat = parallel.all_reduce_sum_host(np.concatenate([s1, s2.reshape(-1), [float(n)]]))
            D = self.ndims
            s1, s2, n = flat[:D], flat[D:D + D * D].reshape(D, D), int(round(flat[-1]))
        # quirk Q3: raw scatter about mu_d, (X - mu)^T (X - mu) = S2 - mu s1^T - s1 mu^T + n mu mu^T, for all N
        # proposals at once (the reference loops over proposals and samples, m_pmc.py:62-72)
        with np.errstate(invalid="ignore", over="ignore"):
            covs = (s2[None, :, :] - mu[:, :, None] * s1[None, None, :] - s1[None, :, None] * mu[:, None, :]
                    + n * mu[:, :, None] * mu[:, None, :])
            bad = np.isnan(covs).any(axis=(1, 2)) | (covs > 10).any(axis=(1, 2))
        covs[bad] = np.diag(np.ones(self.ndims) * 10)
        self._finish_adapt(alpha, mu, covs)

    def _paper_covariances(self, mom, alpha, mu):
        """Sigma_d = sum_i w~_i rho_di (x_i - mu_d)(x_i - mu_d)^T / alpha_d = S2_d / alpha_d - mu_d mu_d^T (eq. 14.3 of
        the M-PMC paper), from the device-side weighted second moments; same failsafe as the reference (NaN or an
        entry > 10 -> 10 I), plus: a component that received no mass keeps its old covariance."""
        D = self.ndims
        tri = mom[:, D + 1:]
        r, c = np.tril_indices(D)
        s2 = np.zeros((self.N, D, D))
        s2[:, r, c] = tri
        s2[:, c, r] = tri
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            covs = s2 / alpha[:, None, None] - mu[:, :, None] * mu[:, None, :]
        dead = ~(alpha > 0)
        covs[dead] = self._covs[dead]
        covs += 1e-9 * np.eye(D)[None]                     # float32 partial sums: keep the factorisation well-posed
        bad = np.isnan(covs).any(axis=(1, 2)) | (covs > 10).any(axis=(1, 2))
        bad |= ~(np.linalg.det(np.where(np.isnan(covs), 0.0, covs)) > 0)
        covs[bad] = np.diag(np.ones(D) * 10)
        return covs

    def _finish_adapt(self, alpha, mu, covs):
        mu = np.where(np.isnan(mu), self._means, mu) if self.paper_covariance else mu
        # CMultivariateNormal.set_moments asserts det > 0 (CMultivariateNormal.py:55-56)
        assert np.all(np.linalg.det(covs) > 0)
        self._means, self._covs = mu, covs
        self.wproposals = alpha
        self.wproposals = self.wproposals / self.wproposals.sum()
        self._mix_weights = sanitize_weights(self.wproposals.copy())
        self._mixture = None
        self._objects = None
        self._sampler_params = None

    def importance_sample(self, target_d, n_samples, timeout=60):
        elapsed_time = 0
        t_ini = time.time()
        # (sharded: the loop is driven by the sample count alone -- a per-rank clock must not desynchronise the collectives)
        while n_samples > (self._n_all_ranks if self.sharded else self._n()) and (self.sharded or elapsed_time < timeout):
            new_samples = self.sample_batch()
            k = int(new_samples.shape[0])
            logw, logq, partials = self.weigh(new_samples, target_d)
            self._num_pi_evals += k
            self._num_q_evals += k * (self.N + 1)
            if self._device_update_ok():
                rec = partials
                if self.sharded:
                    from .. import parallel                 # the batch is the union of all ranks' draws
                    rec = _device.merge_weight_records(parallel.all_gather_records_dev(partials))
                self.adapt_device(new_samples, logw, logq, rec, int(self.K))
                # accumulated weights are the batch-normalised ones (quirk Q4); the normaliser never visits the host
                self._append(new_samples, logw - (rec[0] + torch.log(rec[1])).to(torch.float32))
                self._n_all_ranks += int(self.K)
                elapsed_time = time.time() - t_ini
                continue
            ph = partials.cpu().numpy()
            if self.sharded:
                from .. import parallel
                ph = parallel.global_weight_partials(ph)   # the batch is the union of all ranks' draws
            self.adapt(new_samples, logw, logq, ph)
            # accumulated weights are the batch-normalised ones (quirk Q4)
            self._append(new_samples, logw - float(_device.weight_logsum(ph)))
            self._n_all_ranks += int(self.K)
            elapsed_time = time.time() - t_ini
        return self._out(self._samples_dev), self._out(self._logw_dev, exp=True)

    def sample_batch(self):
        k = int(self.K)
        if self.sharded:
            from .. import parallel
            a, b = parallel.shard_rows(k)
            k = b - a
        self._num_q_samples += k
        return self._sample_device(k)

    def _update_model(self):
        pass
