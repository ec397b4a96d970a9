"""KL and JS divergence metrics with the reference's interface
(reference ais_benchmarks/metrics/divergences.py:59-195).

compute(p=, q=, nsamples=) draws uniform evaluation points on p's support, evaluates both log
densities on the GPU (for q built by a sampler's KDE this is the pairwise N_eval x N_kde kernel)
and reduces them to the self-normalised discrete KL with ONE device pass
(C ABI aisb_kld_partials_f32 -> 8 doubles -> aisb_kld_finalize):

    I   = { i : lp_i > -inf and lq_i > -inf }
    KLD = sum_I exp(lp_i - Lp) (lp_i - lq_i) - Lp + Lq,   Lp = LSE_I lp,  Lq = LSE_I lq

which equals the reference's normalise-in-log-space-then-sum (divergences.py:136-153) exactly.

The JSD (divergences.py:156-195) reuses the same two device vectors: one pass for the inlier normalisers
(aisb_pair_partials_f32 with the float64 underflow floor, the log-domain image of the reference's `prob > 0`),
one pass for the terms (aisb_jsd_terms_f32). With `sharded = True` every rank evaluates its own rows and only
the 8- and 2-double records travel (parallel.py).
"""
import numpy as np
import torch

from .. import _device
from .base import CMetric


def _is_ours(obj):
    """True if the `log_prob` this object answers with is one of this package's (device tensors in, device tensors
    out): decided by the class that DEFINES it, so a user's subclass of a sampler or distribution (the reference's
    own extension mechanism) stays on the device unless it overrides log_prob itself."""
    pkg = __package__.rsplit(".", 1)[0]
    for c in type(obj).__mro__:
        if "log_prob" in c.__dict__:
            return c.__module__.startswith(pkg)
    return False


def _log_prob_device(dist, x_dev, x_np_getter):
    """log density on the device: stays in HBM for this package's classes, goes through numpy for foreign ones."""
    if _is_ours(dist):
        lp = dist.log_prob(x_dev)
        if _device.is_tensor(lp):
            return lp.to(torch.float32)
        return torch.from_numpy(np.asarray(lp, dtype=np.float32)).to(x_dev.device)
    lp = np.asarray(dist.log_prob(x_np_getter()), dtype=np.float64).reshape(-1)
    return torch.from_numpy(lp).to(x_dev.device, torch.float32)


def evaluation_points(metric, kwargs):
    """The argument contract and the uniform evaluation points of CDivergence.compute / CExpectedValueMSE.compute
    (divergences.py:77-87, statistics.py:29-37)."""
    assert "p" in kwargs.keys() and hasattr(kwargs["p"], "log_prob"), \
        "p must implement a log_prob method. p is of type %s" % (type(kwargs["p"]))
    assert "q" in kwargs.keys() and hasattr(kwargs["q"], "log_prob"), \
        "q must implement a log_prob method. q is of type %s" % (type(kwargs["q"]))
    assert "nsamples" in kwargs.keys()
    p = kwargs["p"]
    lo, hi = p.support()[0], p.support()[1]
    if metric.device_points:
        _device.require_cuda()
        dev = _device.device()
        return _device.uniform_box(torch.from_numpy(np.asarray(lo, dtype=np.float32)).to(dev),
                                   torch.from_numpy(np.asarray(hi, dtype=np.float32)).to(dev), int(kwargs["nsamples"]))
    # same host RNG call as divergences.py:84-85, so np.random.seed reproduces the evaluation points
    return np.random.uniform(lo, hi, size=(kwargs["nsamples"], p.dims))


def device_log_probs(p, q, samples, deal_out=False):
    """(x, lp, lq) as float32 device tensors for the evaluation points `samples` (numpy or CUDA tensor).
    deal_out: `samples` is the FULL point set, identical on every rank of a multi-GPU job: every rank stages 1/N of it
    (parallel.upload_replicated), the ranks exchange the pieces over NVLink and each keeps its block-cyclic share of the
    Z-order curve (parallel.spatial_shard) -- the sharding the KDE kernels are fastest on (a rank's warps then hold
    neighbours of the full set; a random 1/8 of the points costs 5 % more per pair)."""
    dims = getattr(p, "dims", None) or samples.shape[-1]
    if deal_out:
        from .. import parallel
        if _device.is_tensor(samples):
            full, was_numpy = _device.to_device_points(samples, dims)[0], False
        else:
            _device.require_cuda()
            full, was_numpy = parallel.upload_replicated(_device.check_points(np.asarray(samples), dims)), True
        x_dev, mine = parallel.spatial_shard(full)
    else:
        x_dev, was_numpy = _device.to_device_points(samples, dims)
        mine = None
    cache = {}

    def x_np():
        # host copy of THIS rank's points, only built if a foreign distribution asks for numpy input
        if "x" not in cache:
            if not was_numpy:
                cache["x"] = _device.download(x_dev)
            else:
                cache["x"] = samples if mine is None else np.asarray(samples)[mine.cpu().numpy()]
        return cache["x"]

    return x_dev, _log_prob_device(p, x_dev, x_np), _log_prob_device(q, x_dev, x_np)


class CDivergence(CMetric):
    def __init__(self):
        super().__init__()
        self.name = "base"
        self.type = "divergence"
        self.is_symmetric = False
        self.disjoint_support = True
        # True: evaluation points are drawn by the device Philox kernel instead of np.random.uniform
        self.device_points = False
        # multi-GPU (with `sharded`): compute_from_samples is handed the FULL point set, identical on every rank, and
        # deals it out itself (device_log_probs, deal_out); False: the caller passes this rank's rows
        self.replicated_points = False
        # True (under torch.distributed): `samples` are this rank's rows; the partial records are merged over the ranks
        self.sharded = False

    def pre(self, **kwargs):
        pass

    def post(self, **kwargs):
        pass

    def reset(self, **kwargs):
        pass

    def compute(self, **kwargs):
        return self.compute_from_samples(kwargs.get("p"), kwargs.get("q"), evaluation_points(self, kwargs))

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
        _, lp, lq = device_log_probs(p, q, samples, deal_out=self.sharded and self.replicated_points)
        return self.compute_from_device_log_probs(lp, lq)

    def compute_from_device_log_probs(self, lp, lq):
        """The reduction alone, on float32 device vectors (shared with the fused evaluation loop)."""
        self.partials = _device.kld_partials(lp, lq).cpu().numpy()
        if self.sharded:
            from .. import parallel
            self.partials = parallel.merge_kld_partials(parallel.all_gather_records(self.partials))
        self._warn(self.partials)
        # np.float64 like the reference's np.sum(res) (divergences.py:105): callers divide by expected values that may
        # be 0 (the reference's own test_equal) and rely on numpy's inf / nan instead of a ZeroDivisionError
        self.value = np.float64(_device.kld_finalize(self.partials))
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


class CJSDivergence(CDivergence):
    """Jensen-Shannon divergence in bits (divergences.py:156-195), both densities and both reductions on the device."""

    def __init__(self):
        super().__init__()
        self.name = "JSD"
        self.type = "divergence"
        self.is_symmetric = True
        self.disjoint_support = False
        self.log2 = np.log(2)
        self.partials = None
        self.dropped = 0

    def compute_from_samples(self, p, q, samples):
        _, lp, lq = device_log_probs(p, q, samples, deal_out=self.sharded and self.replicated_points)
        return self.compute_from_device_log_probs(lp, lq)

    def compute_from_device_log_probs(self, lp, lq):
        from .. import parallel
        floor = _device._lib.LOG_TINY  # the reference filters on linear-domain prob > 0 (:170-173)
        part = _device.pair_partials(lp, lq, floor).cpu().numpy()
        if self.sharded:
            part = parallel.merge_kld_partials(parallel.all_gather_records(part))
        self.partials = part
        if part[7] > 0 and part[5] / part[7] < .1:
            print("WARNING. JSD compute. Only %6.3f%% inliers" % (part[5] / part[7] * 100))
        if part[5] <= 0:
            self.value = np.float64(0.0)  # empty sums (:188-195)
            return self.value
        log_norm_p, log_norm_q = _device.pair_log_norms(part)
        terms = _device.jsd_terms(lp, lq, floor, log_norm_p, log_norm_q)
        terms = parallel.all_gather_records(terms).sum(axis=0) if self.sharded else terms.cpu().numpy()
        self.dropped = int(terms[1])
        self.value = np.float64(terms[0] / self.log2)
        return self.value
