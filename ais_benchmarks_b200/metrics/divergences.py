"""KL divergence metric with the reference's interface
(reference ais_benchmarks/metrics/divergences.py:59-153).

compute(p=, q=, nsamples=) draws uniform evaluation points on p's support, evaluates both log
densities on the GPU (for q built by a sampler's KDE this is the pairwise N_eval x N_kde kernel)
and reduces them to the self-normalised discrete KL with ONE device pass
(C ABI aisb_kld_partials_f32 -> 8 doubles -> aisb_kld_finalize):

    I   = { i : lp_i > -inf and lq_i > -inf }
    KLD = sum_I exp(lp_i - Lp) (lp_i - lq_i) - Lp + Lq,   Lp = LSE_I lp,  Lq = LSE_I lq

which equals the reference's normalise-in-log-space-then-sum (divergences.py:136-153) exactly.
"""
import numpy as np
import torch

from .. import _device
from .base import CMetric


def _is_ours(obj):
    return type(obj).__module__.startswith(__package__.rsplit(".", 1)[0])


def _log_prob_device(dist, x_dev, x_np_getter):
    """log density on the device: stays in HBM for this package's classes, goes through numpy for foreign ones."""
    if _is_ours(dist):
        lp = dist.log_prob(x_dev)
        if _device.is_tensor(lp):
            return lp.to(torch.float32)
        return torch.from_numpy(np.asarray(lp, dtype=np.float32)).to(x_dev.device)
    lp = np.asarray(dist.log_prob(x_np_getter()), dtype=np.float64).reshape(-1)
    return torch.from_numpy(lp).to(x_dev.device, torch.float32)


class CDivergence(CMetric):
    def __init__(self):
        super().__init__()
        self.name = "base"
        self.type = "divergence"
        self.is_symmetric = False
        self.disjoint_support = True
        # True: evaluation points are drawn by the device Philox kernel instead of np.random.uniform
        self.device_points = False

    def pre(self, **kwargs):
        pass

    def post(self, **kwargs):
        pass

    def reset(self, **kwargs):
        pass

    def compute(self, **kwargs):
        assert "p" in kwargs.keys() and hasattr(kwargs["p"], "log_prob"), \
            "p must implement a log_prob method. p is of type %s" % (type(kwargs["p"]))
        assert "q" in kwargs.keys() and hasattr(kwargs["q"], "log_prob"), \
            "q must implement a log_prob method. q is of type %s" % (type(kwargs["q"]))
        assert "nsamples" in kwargs.keys()
        p = kwargs["p"]
        lo, hi = p.support()[0], p.support()[1]
        if self.device_points:
            _device.require_cuda()
            dev = _device.device()
            samples = _device.uniform_box(torch.from_numpy(np.asarray(lo, dtype=np.float32)).to(dev),
                                          torch.from_numpy(np.asarray(hi, dtype=np.float32)).to(dev),
                                          int(kwargs["nsamples"]))
        else:
            # same host RNG call as divergences.py:84-85, so np.random.seed reproduces the evaluation points
            samples = np.random.uniform(lo, hi, size=(kwargs["nsamples"], p.dims))
        return self.compute_from_samples(p, kwargs["q"], samples)

    def compute_from_samples(self, p, q, samples):
        raise NotImplementedError


class CKLDivergence(CDivergence):
    def __init__(self):
        super().__init__()
        self.name = "KL"
        self.type = "divergence"
        self.is_symmetric = False
        self.disjoint_support = False
        self.partials = None

    def compute_from_samples(self, p, q, samples):
        """divergences.py:101-108 with both log densities and the reduction on the device."""
        dims = getattr(p, "dims", None) or samples.shape[-1]
        x_dev, was_numpy = _device.to_device_points(samples, dims)
        cache = {}

        def x_np():
            if "x" not in cache:
                cache["x"] = samples if was_numpy else x_dev.cpu().numpy().astype(np.float64)
            return cache["x"]

        lp = _log_prob_device(p, x_dev, x_np)
        lq = _log_prob_device(q, x_dev, x_np)
        self.partials = _device.kld_partials(lp, lq).cpu().numpy()
        self._warn(self.partials)
        self.value = _device.kld_finalize(self.partials)
        return self.value

    @staticmethod
    def _warn(partials):
        if partials[7] > 0 and partials[5] / partials[7] < .1:
            print("WARNING. KLDD compute. Less than 10% inliers")

    @staticmethod
    def compute_from_probs(p_samples_prob, q_samples_prob):
        """Linear-domain variant (divergences.py:110-133); unused by compute(). Host arrays in, host array out:
        this is API-parity glue, not part of the accelerated path."""
        inliers = np.logical_and(p_samples_prob > 0, q_samples_prob > 0)
        num_inliers = len(p_samples_prob[inliers])
        if num_inliers / len(p_samples_prob) < .1:
            print("WARNING. JSD compute. Less than 10% inliers")
        if np.sum(p_samples_prob[inliers]) > 0:
            p_samples_prob[inliers] = p_samples_prob[inliers] / np.sum(p_samples_prob[inliers])
        if np.sum(q_samples_prob[inliers]) > 0:
            q_samples_prob[inliers] = q_samples_prob[inliers] / np.sum(q_samples_prob[inliers])
        else:
            return np.inf
        res = p_samples_prob[inliers] * np.log(p_samples_prob[inliers] / q_samples_prob[inliers])
        res[q_samples_prob[inliers] <= 0] = 0
        return res

    @staticmethod
    def compute_from_log_probs(p_samples_logprob, q_samples_logprob):
        """Per-inlier KL terms (divergences.py:135-153); their sum is the KLD. The normalisers Lp, Lq come from the
        device reduction; like the reference, numpy inputs are normalised IN PLACE on their inlier entries (quirk Q7)."""
        _device.require_cuda()
        dev = _device.device()
        as_t = _device.is_tensor(p_samples_logprob)
        lp = (p_samples_logprob if as_t else torch.from_numpy(np.asarray(p_samples_logprob))).to(dev, torch.float32)
        lq = (q_samples_logprob if _device.is_tensor(q_samples_logprob)
              else torch.from_numpy(np.asarray(q_samples_logprob))).to(dev, torch.float32)
        part = _device.kld_partials(lp.contiguous(), lq.contiguous()).cpu().numpy()
        CKLDivergence._warn(part)
        with np.errstate(divide="ignore", invalid="ignore"):
            Lp = part[0] + np.log(part[1])
            Lq = part[3] + np.log(part[4])
        inl = (lp > -np.inf) & (lq > -np.inf)
        pn = lp.double()[inl] - Lp
        qn = lq.double()[inl] - Lq
        res = torch.exp(pn) * (pn - qn)
        res[torch.isnan(qn)] = np.inf
        if not as_t:
            mask = inl.cpu().numpy()
            if isinstance(p_samples_logprob, np.ndarray) and p_samples_logprob.dtype.kind == "f":
                p_samples_logprob[mask] -= Lp
            if isinstance(q_samples_logprob, np.ndarray) and q_samples_logprob.dtype.kind == "f":
                q_samples_logprob[mask] -= Lq
            return res.cpu().numpy()
        return res
