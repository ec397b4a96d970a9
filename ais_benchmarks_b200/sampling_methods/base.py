"""Sampling-method base classes with the reference's interface
(reference ais_benchmarks/sampling_methods/base.py:11-246).

State lives on the GPU: accumulated samples as a float32 [n, D] tensor and accumulated
*log* weights as a float32 [n] tensor. ``self.samples`` / ``self.weights`` are numpy views of
that state (float64, weights in the linear domain like the reference) materialised on demand, so
unmodified callers (CBenchmark.evaluate_method, write_results) keep working.
"""
import math
from abc import ABCMeta, abstractmethod

import numpy as np
import torch

from .. import _device
from ..distributions.distributions import CDistribution


class CSamplingMethod(metaclass=ABCMeta):
    def __init__(self, params):
        self.space_min = np.asarray(params["space_min"], dtype=np.float64)
        self.space_max = np.asarray(params["space_max"], dtype=np.float64)
        self.ndims = params["dims"]
        self._name = params["name"] if "name" in params.keys() else "GenericSamplingMethod"
        # True: importance_sample()/sample() return CUDA tensors instead of numpy arrays
        self.device_outputs = bool(params.get("device_outputs", False))
        # multi-GPU SPMD callers: `samples` assigned through the numpy setter are identical on every rank (see setter)
        self.replicated_samples = bool(params.get("replicated_samples", False))

        assert self.space_max.shape == self.space_min.shape
        assert self.ndims == len(self.space_max)

        self._debug = False
        self._reset_state()

    def _reset_state(self):
        self._num_pi_evals = 0
        self._num_q_evals = 0
        self._num_q_samples = 0
        self._samples_dev = None   # float32 [n, D] CUDA
        self._logw_dev = None      # float32 [n] CUDA, log of the reference's self.weights
        self._buf_s = self._buf_w = None  # growth buffers behind the two views (see _append)
        self._samples_np = None
        self._weights_np = None
        self._partials = None      # host float64[4] record of the accumulated log-weights
        self.variogram = np.array([])

    def reset(self):
        self._reset_state()

    # -- numpy views of the device state -----------------------------------------------------------
    @property
    def samples(self):
        if self._samples_np is None:
            self._samples_np = (np.array([]) if self._samples_dev is None
                                else _device.download(self._samples_dev))
        return self._samples_np

    @samples.setter
    def samples(self, val):
        val = np.asarray(val, dtype=np.float64)
        self._samples_np = val
        if val.size:
            _device.require_cuda()
            rows = val.reshape(-1, self.ndims)
            if getattr(self, "replicated_samples", False):
                # SPMD caller's promise: every rank assigns the SAME array -> each stages a slice, NVLink does the rest
                from .. import parallel
                self._samples_dev = parallel.upload_replicated(rows)
            else:
                self._samples_dev = _device.upload_f32(rows)
        else:
            self._samples_dev = None

    @property
    def weights(self):
        if self._weights_np is None:
            self._weights_np = (np.array([]) if self._logw_dev is None
                                else torch.exp(self._logw_dev.double()).cpu().numpy())
        return self._weights_np

    @weights.setter
    def weights(self, val):
        val = np.asarray(val, dtype=np.float64)
        self._weights_np = val
        self._partials = None
        if val.size:
            _device.require_cuda()
            with np.errstate(divide="ignore"):
                self._logw_dev = torch.from_numpy(np.log(val)).to(_device.device(), torch.float32)
        else:
            self._logw_dev = None

    def _append(self, new_samples, new_logw):
        """Accumulate a batch (the reference grows self.samples / self.weights with np.vstack / np.concatenate on
        every iteration, O(n^2) copying: dm_ais.py:69,79-80). Here the state lives in capacity-doubling HBM buffers and
        self._samples_dev / self._logw_dev are prefix views of them, so a run of many small iterations copies O(n)."""
        n0, k = self._n(), int(new_samples.shape[0])
        if n0 == 0:
            # first batch: adopt the tensors as they are (no copy; a 1e7-sample batch must not be duplicated)
            self._buf_s, self._buf_w = new_samples, new_logw
        else:
            owned = (self._buf_s is not None and self._samples_dev.data_ptr() == self._buf_s.data_ptr()
                     and self._logw_dev.data_ptr() == self._buf_w.data_ptr())
            if not owned or n0 + k > self._buf_s.shape[0]:
                cap = max(n0 + k, 2 * n0)
                buf_s = torch.empty((cap, new_samples.shape[1]), dtype=torch.float32, device=new_samples.device)
                buf_w = torch.empty(cap, dtype=torch.float32, device=new_samples.device)
                buf_s[:n0].copy_(self._samples_dev)
                buf_w[:n0].copy_(self._logw_dev)
                self._buf_s, self._buf_w = buf_s, buf_w
            self._buf_s[n0:n0 + k].copy_(new_samples)
            self._buf_w[n0:n0 + k].copy_(new_logw)
        self._samples_dev = self._buf_s[:n0 + k]
        self._logw_dev = self._buf_w[:n0 + k]
        self._samples_np = self._weights_np = self._partials = None

    def _n(self):
        return 0 if self._samples_dev is None else int(self._samples_dev.shape[0])

    def _weight_partials(self):
        if self._partials is None:
            self._partials = _device.logw_partials(self._logw_dev).cpu().numpy()
        return self._partials

    def _out(self, t, exp=False):
        if self.device_outputs:
            return torch.exp(t) if exp else t
        # float64 numpy like the reference; large results are widened on the device (numpy's astype is single-threaded)
        return _device.exp_from_device(t, True) if exp else _device.from_device(t, True)

    # -- statistics --------------------------------------------------------------------------------
    def get_acceptance_rate(self):
        assert self._n(), "Invalid number of samples to compute acceptance rate"
        return self._n() / self._num_q_samples

    def get_NESS(self):
        """Autocorrelation ESS / n of the sample chain (role of base.py:41-70; used by the sequential samplers, the
        importance samplers override it with get_approx_NESS). Lag-t autocorrelation from the variogram,
        rho_t = 1 - V_t / (2 W) with V_t the mean squared lag-t increment and W the sample variance (both summed over
        all coordinates), accumulated until the first even lag at which rho_{t-1} + rho_t turns negative (Geyer's
        initial positive sequence); ESS = n / (1 + 2 sum_t rho_t). Host statistic, not on the hot path."""
        x = np.asarray(self.samples, dtype=np.float64)
        n = len(x)
        if n < 2:
            return 1.0
        two_w = 2.0 * np.sum((x - x.mean()) ** 2) / (n - 1)
        tail, prev = 0.0, 1.0
        for lag in range(1, n):
            step = x[lag:] - x[:n - lag]
            rho = 1.0 - (np.sum(step * step) / (n - lag)) / two_w
            tail += rho
            if lag % 2 == 0 and prev + rho < 0:
                break
            prev = rho
        return 1.0 / (1.0 + 2.0 * tail)

    def get_approx_NESS(self):
        """ESS / n with ESS = 1 / sum w~^2 (base.py:72-75), from the device-side partial sums
        (C ABI aisb_logw_partials_f32 + aisb_weight_partials_ess)."""
        return _device.weight_ess(self._weight_partials()) / self._n()

    def get_stats(self):
        return {"proposal_samples": self._num_q_samples,
                "proposal_evals": self._num_q_evals,
                "target_evals": self._num_pi_evals,
                "n_samples": self._n()}

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, val):
        self._name = val

    @property
    def num_proposal_samples(self):
        return self._num_q_samples

    @property
    def num_target_evals(self):
        return self._num_pi_evals

    @property
    def num_proposal_evals(self):
        return self._num_q_evals

    @property
    def debug(self):
        return self._debug

    @debug.setter
    def debug(self, val):
        self._debug = val

    @abstractmethod
    def sample(self, n_samples):
        raise NotImplementedError

    @abstractmethod
    def importance_sample(self, target_d, n_samples, timeout):
        raise NotImplementedError

    @abstractmethod
    def prob(self, samples):
        raise NotImplementedError

    @abstractmethod
    def log_prob(self, samples):
        raise NotImplementedError

    def draw(self, ax):
        return []

    def get_viz_frames(self):
        return None


class CDeviceKDE(CDistribution):
    """The KDE that CMixtureSamplingMethod._update_model builds (base.py:162-179), kept as ONE packed
    HBM buffer of n_kde centres instead of n_kde CMultivariateNormal objects: equal weights, kernel
    covariance bw*I (bw is a variance, quirk Q1)."""

    def __init__(self, samples_dev, idx_dev, n_kde, bw, dims, support):
        params = {"type": "KDE", "family": "mixture", "dims": dims, "support": support,
                  "likelihood_f": self.prob, "loglikelihood_f": self.log_prob}
        super().__init__(params)
        self.bw = float(bw)
        self.n_kde = int(n_kde)
        self._idx = idx_dev
        self._src = samples_dev
        self.mixture = _device.DeviceMixture.kde(samples_dev, idx_dev, n_kde, bw)

    @property
    def weights(self):
        """Equal mixture weights 1 / n_kde (base.py:177), materialised only when somebody reads them: filling a
        1e6-element numpy array costs ~1 ms of host time, more than the whole device-side build of the KDE."""
        return np.full(shape=(self.n_kde,), fill_value=1 / self.n_kde)

    def centres(self):
        return self._src if self._idx is None else self._src[self._idx]

    def log_prob(self, data):
        x, as_numpy = _device.to_device_points(data, self.dims)
        return _device.from_device(_device.mixture_logpdf(x, self.mixture), as_numpy)

    def prob(self, data):
        x, as_numpy = _device.to_device_points(data, self.dims)
        return _device.exp_from_device(_device.mixture_logpdf(x, self.mixture), as_numpy)

    def sample(self, n_samples=1, as_tensor=False):
        c = self.centres()
        pick = torch.randint(0, c.shape[0], (int(n_samples),), device=c.device)
        x = _device.sample_iso(c[pick].contiguous(), 1, self.bw)
        return x if as_tensor else _device.download(x)

    def condition(self, dist):
        raise NotImplementedError

    def marginal(self, dim):
        raise NotImplementedError


class CMixtureSamplingMethod(CSamplingMethod):
    def __init__(self, params):
        super(CMixtureSamplingMethod, self).__init__(params)
        self.model = None
        self.n_samples_kde = params["n_samples_kde"]

    def reset(self):
        super(CMixtureSamplingMethod, self).reset()
        self.model = None

    def sample(self, n_samples):
        if self.model is None:
            self._update_model()
        self._num_q_samples += n_samples
        return self.model.sample(n_samples, as_tensor=True) if self.device_outputs else self.model.sample(n_samples)

    def prob(self, s):
        if self.model is None:
            self._update_model()
        return self.model.prob(s)

    def log_prob(self, s):
        if self.model is None:
            self._update_model()
        return self.model.log_prob(s)

    def _update_model(self):
        """KDE over n_samples_kde samples drawn with replacement (base.py:162-179)."""
        _device.require_cuda()
        n = self._n()
        assert n > 0, "no samples to build the KDE from"
        seed = int(np.random.randint(0, 2 ** 31 - 1))  # follows np.random.seed like the reference's index draw
        gen = torch.Generator(device=_device.device())
        gen.manual_seed(seed)
        idx = torch.randint(0, n, (int(self.n_samples_kde),), device=_device.device(), generator=gen)
        self.model = CDeviceKDE(self._samples_dev, idx, self.n_samples_kde, self.bw, self.ndims,
                                [self.space_min, self.space_max])


class CMixtureISSamplingMethod(CMixtureSamplingMethod):
    def __init__(self, params):
        params["n_samples_kde"] = None
        super(CMixtureISSamplingMethod, self).__init__(params)

    def get_NESS(self):
        return self.get_approx_NESS()

    def _update_model(self):
        # reference base.py:234-246 is broken as shipped (wrong CMixtureModel call) and is overridden with `pass`
        # by every importance sampler on the path; keep the override contract.
        raise NotImplementedError("CMixtureISSamplingMethod subclasses own their proposal mixture")


# ---------------------------------------------------------------------------------------------------
# helpers shared by the importance samplers
# ---------------------------------------------------------------------------------------------------
def target_device_mixture(target_d):
    """Packed HBM parameters of a target built from this package's Gaussian classes, else None."""
    if hasattr(target_d, "device_mixture"):
        return target_d.device_mixture()
    if hasattr(target_d, "_packed") and callable(getattr(target_d, "_packed")):
        kind, packed = target_d._packed()
        return packed if kind == "gauss" else None
    if hasattr(target_d, "_device_mixture"):
        return target_d._device_mixture()
    return None


def foreign_target_logp(target_d, x_dev):
    """log pi(x) of a target that is not one of this package's Gaussian classes: the caller's own log_prob/prob
    runs wherever it runs (typically numpy on the host); the result is uploaded for the weight kernel."""
    x_np = _device.download(x_dev)
    if hasattr(target_d, "log_prob"):
        lp = np.asarray(target_d.log_prob(x_np), dtype=np.float64).reshape(-1)
    else:
        with np.errstate(divide="ignore"):
            lp = np.log(np.asarray(target_d.prob(x_np), dtype=np.float64).reshape(-1))
    return torch.from_numpy(lp).to(x_dev.device, torch.float32)


mixture_component_draws = _device.mixture_component_draws


def multinomial_resample(logw, n_draws):
    """n_draws indices ~ Categorical(w / sum w), w = exp(logw) (dm_ais.py:47-54). Invalid (NaN / +inf) weights
    count as zero; if no weight is left the draw is uniform. Inverse-CDF on the device: cumsum + searchsorted, float64.
    No host synchronisation (the degenerate case is a device-side select), so the call can sit in a CUDA graph."""
    lw = logw.double()
    lw = torch.where(torch.isfinite(lw) | (lw == -math.inf), lw, torch.full_like(lw, -math.inf))
    m = torch.max(lw)
    ok = torch.isfinite(m)
    w = torch.where(ok, torch.exp(lw - torch.where(ok, m, torch.zeros_like(m))), torch.ones_like(lw))
    cdf = torch.cumsum(w, 0)
    u = torch.rand(n_draws, dtype=torch.float64, device=logw.device) * cdf[-1]
    return torch.clamp(torch.searchsorted(cdf, u, right=True), max=logw.shape[0] - 1)
