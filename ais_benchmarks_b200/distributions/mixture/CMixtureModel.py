"""Weighted mixture over a list of component distributions, reference API
(reference ais_benchmarks/distributions/mixture/CMixtureModel.py:8-89).

The reference loops over components in Python, materialises a (K, N) float64 matrix and calls
scipy logsumexp(b=weights). Here the whole mixture is packed once into HBM (one row per
component) and evaluated by a single sm_100a kernel with an online log-sum-exp
(C ABI aisb_mixture_logpdf_f32 / aisb_box_mixture_logpdf_f32).

Supported component sets: all Gaussian (CMultivariateNormal, any covariance) or all box-uniform
(CMultivariateUniform). Other / mixed component types are outside the accelerated path
(DESIGN.md "Out of scope") and raise NotImplementedError instead of silently running on the CPU.
"""
import sys

import numpy as np
import torch

from ... import _device
from ..._lib import COV_DIAG, COV_FULL, COV_ISO_SHARED
from ..distributions import CDistribution


def _is_normal(m):
    return getattr(m, "type", None) == "normal" and hasattr(m, "loc") and hasattr(m, "scale")


def _is_uniform(m):
    return getattr(m, "type", None) == "uniform" and hasattr(m, "loc") and hasattr(m, "scale")


def sanitize_weights(weights):
    """CMixtureModel.set_weights (CMixtureModel.py:25-41): NaN / <= 0 -> 0 (in place, with the reference's stderr
    warning), then divide by the sum when it is positive."""
    nan_indices = np.logical_or(np.isnan(weights), weights <= 0)
    if np.any(nan_indices):
        print(f"CMixtureModel. WARNING! There are {np.count_nonzero(nan_indices)} NaN, negative or "
              f"zero weights out of {len(weights)}. ", file=sys.stderr)
        weights[nan_indices] = 0
    if weights.sum() <= 0:
        print("CMixtureModel. WARNING! All weights are set to zero!. "
              "Using equal weights for all components.", file=sys.stderr)
    else:
        weights = weights / weights.sum()
    return weights


def gaussian_pack_args(models):
    """Component list -> (means [K, D], cov, kind) choosing the cheapest packed layout that is exact."""
    K = len(models)
    D = len(np.asarray(models[0].loc).reshape(-1))
    means = np.empty((K, D), dtype=np.float64)
    covs = np.empty((K, D, D), dtype=np.float64)
    for k, m in enumerate(models):
        means[k] = np.asarray(m.loc, dtype=np.float64).reshape(-1)
        covs[k] = np.asarray(m.scale, dtype=np.float64)
    diag = np.diagonal(covs, axis1=1, axis2=2)
    off = covs - diag[:, :, None] * np.eye(D)[None]
    if np.count_nonzero(off) != 0:
        return means, covs, COV_FULL
    if np.all(diag == diag[0, 0]):
        return means, np.array([diag[0, 0]]), COV_ISO_SHARED
    return means, np.ascontiguousarray(diag), COV_DIAG


class CMixtureModel(CDistribution):
    def __init__(self, params):
        self._check_param(params, "models")
        self._check_param(params, "weights")

        params["type"] = "mixture"
        params["family"] = "mixture"
        params["likelihood_f"] = self.prob
        params["loglikelihood_f"] = self.log_prob
        super(CMixtureModel, self).__init__(params)

        models = params["models"]
        weights = params["weights"]
        assert len(models) == len(weights), "len(models)=%d , len(weights)=%d" % (len(models), len(weights))
        self.models = models
        self.weights = weights
        self._cache_key = None
        self._cache = None
        self.set_weights(np.array(weights, dtype=np.float64))

    def set_weights(self, weights):
        assert len(self.models) == len(weights), "len(models)=%d , len(weights)=%d" % (len(self.models), len(weights))
        self.weights = sanitize_weights(weights)
        self._cache_key = None

    # -- packing ---------------------------------------------------------------------------------
    def _key(self):
        vers = []
        for m in self.models:
            v = getattr(m, "_version", None)
            if v is None:
                return None  # foreign component objects: no change tracking, repack on every call
            vers.append((id(m), v))
        return (tuple(vers), self.weights.tobytes())

    def _packed(self):
        key = self._key()
        if key is not None and key == self._cache_key:
            return self._cache
        if all(_is_normal(m) for m in self.models):
            means, cov, kind = gaussian_pack_args(self.models)
            packed = ("gauss", _device.DeviceMixture.from_params(means, cov, self.weights, kind))
        elif all(_is_uniform(m) for m in self.models):
            packed = ("box", self._pack_boxes(list(range(len(self.models)))))
        else:
            # heterogeneous component list (CMixtureModel.py:53-56 accepts any objects with log_prob): the Gaussian
            # components form one packed sub-mixture, the uniform ones one box mixture, anything else is evaluated
            # through its own log_prob; log_prob joins the groups with a log-sum-exp on the device
            ig = [k for k, m in enumerate(self.models) if _is_normal(m)]
            ib = [k for k, m in enumerate(self.models) if _is_uniform(m)]
            io = [k for k in range(len(self.models)) if k not in ig and k not in ib]
            groups = []
            if ig and self.weights[ig].sum() > 0:
                means, cov, kind = gaussian_pack_args([self.models[k] for k in ig])
                # from_params renormalises the sub-weights: the group's total mass comes back as an offset
                groups.append(("gauss", _device.DeviceMixture.from_params(means, cov, self.weights[ig], kind),
                               float(np.log(self.weights[ig].sum()))))
            if ib:
                groups.append(("box", self._pack_boxes(ib), 0.0))
            packed = ("mixed", (groups, [(self.models[k], float(self.weights[k])) for k in io if self.weights[k] > 0]))
        self._cache_key, self._cache = key, packed
        return packed

    def _pack_boxes(self, which):
        _device.require_cuda()
        dev = _device.device()
        K, D = len(which), self.dims
        lo = np.empty((K, D)); hi = np.empty((K, D)); lwv = np.empty(K)
        for j, k in enumerate(which):
            m = self.models[k]
            c = np.asarray(m.loc, dtype=np.float64).reshape(-1)
            r = np.broadcast_to(np.asarray(m.scale, dtype=np.float64).reshape(-1), (D,))
            lo[j], hi[j] = c - r, c + r
            with np.errstate(divide="ignore"):
                lwv[j] = np.log(self.weights[k]) - np.sum(np.log(2.0 * r))
        return (torch.from_numpy(lo).to(dev, torch.float32), torch.from_numpy(hi).to(dev, torch.float32),
                torch.from_numpy(lwv).to(dev, torch.float32))

    def _points(self, data):
        if not _device.is_tensor(data):
            data = np.asarray(data)
        if len(data.shape) not in (1, 2):
            raise ValueError("Unsupported samples data format: " + str(data.shape))
        return _device.to_device_points(data, self.dims)

    def _log_prob_device(self, x):
        kind, packed = self._packed()
        if kind == "gauss":
            return _device.mixture_logpdf(x, packed)
        if kind == "box":
            lo, hi, lwv = packed
            return _device.box_mixture_logpdf(x, lo, hi, lwv, open_upper=False)
        groups, others = packed
        parts = []
        for gkind, gpacked, goff in groups:
            if gkind == "gauss":
                parts.append(_device.mixture_logpdf(x, gpacked) + goff)
            else:
                parts.append(_device.box_mixture_logpdf(x, gpacked[0], gpacked[1], gpacked[2], open_upper=False))
        for m, w in others:                              # foreign components: their own log_prob, wherever it runs
            lp = m.log_prob(x) if getattr(type(m), "__module__", "").startswith("ais_benchmarks_b200") else \
                m.log_prob(_device.download(x))
            lp = lp if _device.is_tensor(lp) else torch.from_numpy(np.asarray(lp, dtype=np.float64).reshape(-1))
            parts.append(lp.to(x.device, torch.float32) + float(np.log(w)))
        if not parts:
            return torch.full((x.shape[0],), -np.inf, dtype=torch.float32, device=x.device)
        return torch.logsumexp(torch.stack(parts), dim=0)

    def log_prob(self, data):
        """log sum_k w_k p_k(x) (CMixtureModel.py:43-57)."""
        x, as_numpy = self._points(data)
        return _device.from_device(self._log_prob_device(x), as_numpy)

    def prob(self, data):
        """sum_k w_k p_k(x) (CMixtureModel.py:59-77), computed as exp(log_prob) in float64."""
        x, as_numpy = self._points(data)
        return _device.exp_from_device(self._log_prob_device(x), as_numpy)

    def sample(self, n_samples=1, as_tensor=False):
        """Ancestral sampling (CMixtureModel.py:79-89), vectorised on the device: component index ~ weights, then
        one draw from that component."""
        kind, packed = self._packed()
        dev = _device.device()
        n = int(n_samples)
        idx = _device.mixture_component_draws(self.weights, n, dev)
        if kind == "mixed":
            # heterogeneous list: each component draws its own share (like CMixtureModel.py:84-88, batched per component)
            counts = torch.bincount(idx, minlength=len(self.models)).cpu().numpy()
            x = torch.empty((n, self.dims), dtype=torch.float32, device=dev)
            for k, cnt in enumerate(counts):
                if cnt:
                    draw = self.models[k].sample(int(cnt))
                    draw = draw if _device.is_tensor(draw) else torch.from_numpy(np.asarray(draw, dtype=np.float64))
                    x[idx == k] = draw.to(dev, torch.float32).reshape(int(cnt), self.dims)
            return x if as_tensor else _device.download(x)
        if kind == "gauss":
            means, cov, ckind = gaussian_pack_args(self.models)
            z = _device.sample_iso(torch.zeros((1, self.dims), dtype=torch.float32, device=dev), n, 1.0)
            mu = torch.from_numpy(means).to(dev, torch.float32)[idx]
            if ckind == COV_ISO_SHARED:
                x = mu + z * float(np.sqrt(cov[0]))
            elif ckind == COV_DIAG:
                x = mu + z * torch.from_numpy(np.sqrt(cov)).to(dev, torch.float32)[idx]
            else:
                L = np.linalg.cholesky(0.5 * (cov + np.transpose(cov, (0, 2, 1))))
                x = mu + torch.bmm(torch.from_numpy(L).to(dev, torch.float32)[idx], z.unsqueeze(-1)).squeeze(-1)
        else:
            lo, hi, _ = packed
            u = _device.uniform_box(torch.zeros(self.dims, dtype=torch.float32, device=dev),
                                    torch.ones(self.dims, dtype=torch.float32, device=dev), n)
            x = lo[idx] + (hi[idx] - lo[idx]) * u
        return x if as_tensor else _device.download(x)

    def condition(self, dist):
        raise NotImplementedError

    def marginal(self, dim):
        raise NotImplementedError
