"""Tree-pyramid adaptive importance sampling (TP-AIS) with the reference's interface
(reference ais_benchmarks/sampling_methods/tree_pyramid.py:269-741).

What runs where:
  * the 2^D-ary space-partition tree is inherently sequential bookkeeping and stays on the host, here as flat
    numpy arrays indexed by node id (centres, radii, weights, current sample, leaf mask) plus short per-node
    histories, instead of the reference's linked node objects;
  * everything batched goes to the GPU: the target density of all new-node draws of an iteration in one call,
    the target density of ALL leaves in `resample` (tree_pyramid.py:500-533) in one call, and the proposal density
    `prob` / `log_prob` -- for the "haar" kernel one box-mixture kernel launch over the leaves
    (C ABI aisb_box_mixture_logpdf_f32, strict `lower < x < upper` test of tree_pyramid.py:593-596) instead of a
    Python loop over samples x nodes; for the "normal" kernel the Gaussian mixture kernel.

Only ``method="simple"`` is provided: in the reference, "dm" and "mixture" (tree_pyramid.py:363-419) die on
``np.int`` (:395) and, once the root has been split, on the unbounded recursion of ``CTreePyramid.find``
(:140-155), so there is no reference behaviour to match.

Reference quirks kept on purpose: node proposal variance is radius / bw_div for drawing (:55-58) but radius in the
mixture model (:722-723); the "normal" resampling weight divides by the PROTOTYPE density N(0, 0.5 I) without the
Jacobian (:509-511); `_get_samples` with ess_target <= 1 writes leaf i's history at offset i, not at the running
offset (:690-702), and drops zero weights.
"""
import itertools
import time

import numpy as np
import torch

from .. import _device
from ..distributions import CMixtureModel, CMultivariateNormal, CMultivariateUniform
from .base import CMixtureISSamplingMethod


class DeviceDraws:
    """Default source of the random draws TP-AIS needs (device Philox kernels). A test can substitute an object with
    the same two methods to replay draws captured from the reference."""

    def uniform(self, n, dims):
        """n draws ~ U(-0.5, 0.5)^dims, float64 host array."""
        dev = _device.device()
        lo = torch.full((dims,), -0.5, dtype=torch.float32, device=dev)
        hi = torch.full((dims,), 0.5, dtype=torch.float32, device=dev)
        return _device.uniform_box(lo, hi, int(n)).cpu().numpy().astype(np.float64)

    def normal(self, n, dims):
        """n draws ~ N(0, I_dims), float64 host array."""
        dev = _device.device()
        z = _device.sample_iso(torch.zeros((1, dims), dtype=torch.float32, device=dev), int(n), 1.0)
        return z.cpu().numpy().astype(np.float64)


class CTreePyramid:
    """Flat-array 2^D-ary partition tree. Node i: centers[i], radii[i] (half side), weights[i], samples[i],
    is_leaf[i]; `leaves` lists leaf node ids in the reference's leaf-list order (first child takes its parent's slot,
    tree_pyramid.py:209-214), which fixes the tie-breaking of the value sort and the `_get_samples` layout."""

    def __init__(self, dmin, dmax, kernel, bw_div=4):
        dmin, dmax = np.asarray(dmin, dtype=np.float64), np.asarray(dmax, dtype=np.float64)
        if kernel not in ("haar", "normal"):
            raise ValueError("Unknown kernel type. Must be 'normal' or 'haar'")
        self.ndims = len(dmin)
        self.kernel = kernel
        self.bw_div = bw_div
        self.centers = ((dmin + dmax) / 2).reshape(1, -1)
        self.radii = np.array([np.max((dmax - dmin) / 2)])
        self.weights = np.zeros(1)
        self.samples = np.zeros((1, self.ndims))
        self.is_leaf = np.array([True])
        self.level = np.array([0])
        self.node_weight = [0.0]      # the node's own running weight (reference: node.weight)
        self.value = [0.0]            # importance volume used to rank leaves (reference: node.value)
        self.hist_x = [None]          # per-node sample history (reference: node.coords_hist)
        self.hist_w = [None]          # per-node weight history (reference: node.weight_hist)
        self.leaves = [0]

    def __len__(self):
        return len(self.radii)

    @property
    def leaves_idx(self):
        return self.is_leaf

    def record(self, node, x, w):
        """node.sample + node.weigh bookkeeping (tree_pyramid.py:66-88): append to the histories, value = mean(w) r^D."""
        x = np.asarray(x, dtype=np.float64).reshape(1, -1)
        self.hist_x[node] = x if self.hist_x[node] is None else np.concatenate((self.hist_x[node], x))
        self.hist_w[node] = np.array([w]) if self.hist_w[node] is None else np.append(self.hist_w[node], w)
        self.node_weight[node] = w
        self.value[node] = float(np.mean(self.hist_w[node]) * self.radii[node] ** self.ndims)

    def expand(self, node, ess_target=None, n_min=5):
        """Split a leaf into its 2^D children unless the ESS criterion says it is explored well enough
        (tree_pyramid.py:160-235). Returns the new node ids or None."""
        assert self.is_leaf[node], "Non leaf node expansion is not allowed: %d" % node
        if ess_target is not None:
            hw = self.hist_w[node]
            if hw is None or len(hw) < n_min:
                return None
            with np.errstate(divide="ignore", invalid="ignore"):
                ess = (np.sum(hw) * np.sum(hw)) / np.sum(hw * hw)
            if ess > ess_target * len(hw):
                return None
        d = self.ndims
        signs = np.array(list(itertools.product([-1, 1], repeat=d)), dtype=np.float64)
        new_r = self.radii[node] / 2
        kids_c = self.centers[node] + signs * new_r
        nk = len(kids_c)
        first = len(self)
        ids = list(range(first, first + nk))
        self.centers = np.concatenate((self.centers, kids_c))
        self.radii = np.concatenate((self.radii, np.full(nk, new_r)))
        self.weights = np.concatenate((self.weights, np.zeros(nk)))
        self.samples = np.concatenate((self.samples, np.zeros((nk, d))))
        self.is_leaf = np.concatenate((self.is_leaf, np.ones(nk, dtype=bool)))
        self.level = np.concatenate((self.level, np.full(nk, self.level[node] + 1)))
        share = self.node_weight[node] / nk      # children start with their share of the parent's weight (:201-203)
        self.node_weight.extend([share] * nk)
        self.value.extend([float(share * new_r ** d)] * nk)
        self.hist_x.extend([None] * nk)
        self.hist_w.extend([None] * nk)
        self.hist_x[node] = self.hist_w[node] = None
        self.is_leaf[node] = False
        slot = self.leaves.index(node)
        self.leaves[slot] = ids[0]
        self.leaves.extend(ids[1:])
        return ids


class CTreePyramidSampling(CMixtureISSamplingMethod):
    def __init__(self, params):
        """params: space_min, space_max, dims, method ("simple"), resampling ("leaf" | "none"), kernel ("haar" |
        "normal"), ess_target, n_min, parallel_samples -- as tree_pyramid.py:270-316."""
        params["name"] = "TP_AIS"
        super(CTreePyramidSampling, self).__init__(params)
        self.method = params["method"]
        assert self.method in ["simple", "dm", "mixture"], "Invalid method."
        if self.method != "simple":
            raise NotImplementedError("TP-AIS methods 'dm' and 'mixture' cannot run in the reference (np.int at "
                                      "tree_pyramid.py:395, unbounded recursion in CTreePyramid.find :140-155); only "
                                      "'simple' is provided")
        self.ess_target = params["ess_target"]
        self.n_min = params["n_min"]
        self.resampling = params["resampling"]
        assert self.resampling in ["leaf", "none"], "Invalid resampling strategy"
        self.kernel = params["kernel"]
        assert self.kernel in ["normal", "haar"], "Invalid kernel type."
        self.parallel_samples = params["parallel_samples"]
        self.draws = DeviceDraws()
        self.reset()

    def reset(self):
        super(CTreePyramidSampling, self).reset()
        self.T = CTreePyramid(self.space_min, self.space_max, kernel=self.kernel)
        self.model = self._node_sampler(0)
        self._boxes = None

    # -- per-node proposal ("sampling kernel") -------------------------------------------------------------
    def _node_sampler(self, node):
        c, r = self.T.centers[node], self.T.radii[node]
        if self.kernel == "haar":
            return CMultivariateUniform({"center": c.copy(), "radius": np.array([r])})
        return CMultivariateNormal({"mean": c.copy(), "sigma": np.diag(np.ones(self.ndims) * (r / self.T.bw_div))})

    def _draw_nodes(self, nodes):
        """One draw from each listed node's own proposal + that proposal's density at the draw (:66-76)."""
        c, r = self.T.centers[nodes], self.T.radii[nodes].reshape(-1, 1)
        if self.kernel == "haar":
            x = c + 2 * r * self.draws.uniform(len(nodes), self.ndims)
            q = 1.0 / (2 * r.reshape(-1)) ** self.ndims
            # a draw can land exactly on the closed/open boundary only with probability ~0; the reference's
            # CMultivariateUniform.prob would then return 0 (min < x <= max)
        else:
            var = r / self.T.bw_div
            z = self.draws.normal(len(nodes), self.ndims)
            x = c + np.sqrt(var) * z
            q = np.exp(-0.5 * np.sum(z * z, axis=1)) / (2 * np.pi * var.reshape(-1)) ** (self.ndims / 2)
        return x, q

    def _target_prob(self, target_d, x):
        """pi(x) for a batch, float64 host array; on the GPU when the target is one of this package's classes."""
        return np.asarray(target_d.prob(np.ascontiguousarray(x, dtype=np.float64)), dtype=np.float64).reshape(-1)

    def _sample_and_weigh(self, nodes, target_d):
        if not nodes:
            return
        x, q = self._draw_nodes(nodes)
        p = self._target_prob(target_d, x)
        with np.errstate(divide="ignore", invalid="ignore"):
            w = p / q
        for i, node in enumerate(nodes):
            self.T.samples[node] = x[i]
            self.T.record(node, x[i], w[i])
            self.T.weights[node] = w[i]
        self._boxes = None

    # -- the adapt loop ------------------------------------------------------------------------------------
    def importance_sample(self, target_d, n_samples, timeout=60):
        assert self.resampling in ["leaf", "none"], "Unknown resampling strategy"
        elapsed_time = 0
        t_ini = time.time()
        if self.T.hist_w[0] is None and len(self.T) == 1:
            # first call: one draw from the root proposal (:431-438; `simple` counts only the sample here)
            self._sample_and_weigh([0], target_d)
            self._num_q_samples += 1
            self._update_model()
        while n_samples > self._get_nsamples() and elapsed_time < timeout:
            self.step(target_d)
            elapsed_time = time.time() - t_ini
        self._update_model()
        samples, weights = self._get_samples()
        if self.device_outputs:
            dev = _device.device()
            return torch.from_numpy(samples).to(dev, torch.float32), torch.from_numpy(weights).to(dev, torch.float32)
        return samples, weights

    def step(self, target_d):
        """One TP-AIS iteration (:440-454): rank leaves by importance volume, split / re-draw, weigh, resample."""
        order = sorted(self.T.leaves, key=lambda nd: self.T.value[nd], reverse=True)
        new_nodes = self._expand_nodes(order, self.parallel_samples)
        self._sample_and_weigh(new_nodes, target_d)
        self._num_q_samples += len(new_nodes)
        self._num_pi_evals += len(new_nodes)
        self._num_q_evals += len(new_nodes)
        if self.resampling == "leaf":
            self.resample(target_d)

    def _expand_nodes(self, particles, max_parts):
        new_particles = []
        for p in particles:
            kids = self.T.expand(p, self.ess_target, self.n_min)
            if kids is not None:
                new_particles.extend(kids)
                if len(new_particles) > max_parts:
                    return new_particles
            else:
                new_particles.append(p)
        return new_particles

    def resample(self, target_d):
        """Re-draw every leaf's sample from its (scaled prototype) proposal and re-weigh all leaves with ONE batched
        target evaluation (:500-533)."""
        leaf_nodes = np.flatnonzero(self.T.is_leaf)
        n = len(leaf_nodes)
        centers = self.T.centers[leaf_nodes]
        radii = self.T.radii[leaf_nodes].reshape(n, 1)
        if self.kernel == "haar":
            re = self.draws.uniform(n, self.ndims)
            proposal_probs = 1 / ((2 * radii) ** self.ndims)
        else:
            z = self.draws.normal(n, self.ndims)
            re = np.sqrt(0.5) * z                       # prototype N(0, 0.5 I) (:131-133)
            proposal_probs = np.exp(-np.sum(z * z, axis=1) / 2) / (np.pi ** (self.ndims / 2))
        samples = re * 2 * radii + centers
        self.T.samples[leaf_nodes] = samples
        probs = self._target_prob(target_d, samples)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.T.weights[leaf_nodes] = probs / proposal_probs.reshape(-1)
        self._num_pi_evals += n
        self._num_q_evals += n
        self._num_q_samples += n
        self._self_normalize()
        for node in self.T.leaves:
            w = self.T.weights[node]
            self.T.node_weight[node] = w
            if self.T.hist_w[node] is not None:
                self.T.hist_w[node][-1] = w
                self.T.hist_x[node][-1] = self.T.samples[node]
            self.T.value[node] = float(w * (2 * self.T.radii[node]) ** self.ndims)
        self._boxes = None

    # -- proposal density ----------------------------------------------------------------------------------
    def _leaf_boxes(self):
        if self._boxes is None:
            _device.require_cuda()
            dev = _device.device()
            leaf_nodes = np.flatnonzero(self.T.is_leaf)
            c, r = self.T.centers[leaf_nodes], self.T.radii[leaf_nodes].reshape(-1, 1)
            with np.errstate(divide="ignore"):
                lwv = np.log(self.T.weights[leaf_nodes]) - self.ndims * np.log(2 * r.reshape(-1))
            lwv = np.where(np.isnan(lwv), -np.inf, lwv)
            self._boxes = (torch.from_numpy(c - r).to(dev, torch.float32), torch.from_numpy(c + r).to(dev, torch.float32),
                           torch.from_numpy(lwv).to(dev, torch.float32))
        return self._boxes

    def _check(self, x):
        if not _device.is_tensor(x):
            x = np.asarray(x)
        if len(x.shape) not in (1, 2):
            raise ValueError("Unsupported samples data format: " + str(x.shape))
        return x

    def log_prob(self, x):
        x = self._check(x)
        if self.kernel == "normal":
            return super().log_prob(x)
        xd, as_numpy = _device.to_device_points(x, self.ndims)
        lo, hi, lwv = self._leaf_boxes()
        return _device.from_device(_device.box_mixture_logpdf(xd, lo, hi, lwv, open_upper=True), as_numpy)

    def prob(self, x):
        x = self._check(x)
        if self.kernel == "normal":
            return super().prob(x)
        xd, as_numpy = _device.to_device_points(x, self.ndims)
        lo, hi, lwv = self._leaf_boxes()
        return _device.exp_from_device(_device.box_mixture_logpdf(xd, lo, hi, lwv, open_upper=True), as_numpy)

    # -- statistics / outputs ------------------------------------------------------------------------------
    def get_stats(self):
        return {"proposal_samples": self._num_q_samples, "proposal_evals": self._num_q_evals,
                "target_evals": self._num_pi_evals, "n_samples": self._get_nsamples(), "NESS": self.get_approx_NESS()}

    def get_approx_NESS(self):
        return self.get_NESS()

    def get_acceptance_rate(self):
        return self._get_nsamples() / self._num_q_samples

    def get_NESS(self):
        samples, weights = self._get_samples()
        ess = (np.sum(weights) * np.sum(weights)) / np.sum(weights * weights)
        return ess / len(samples)

    def _get_nsamples(self):
        return len(self._get_samples()[1])

    def _get_samples(self):
        if self.ess_target is not None and self.ess_target > 1.0:
            return self.T.samples[self.T.is_leaf], self.T.weights[self.T.is_leaf]
        counts = [0 if self.T.hist_w[nd] is None else len(self.T.hist_w[nd]) for nd in self.T.leaves]
        total = int(np.sum(counts))
        samples = np.zeros((total, self.ndims))
        weights = np.zeros(total)
        for i, nd in enumerate(self.T.leaves):   # reference quirk: leaf i's history lands at offset i (:697-701)
            k = counts[i]
            if k:
                samples[i:i + k] = self.T.hist_x[nd]
                weights[i:i + k] = self.T.hist_w[nd].flatten()
        keep = weights > 0
        return samples[keep], weights[keep]

    @property
    def samples(self):
        return self._get_samples()[0]

    @samples.setter
    def samples(self, val):
        pass

    @property
    def weights(self):
        return self._get_samples()[1]

    @weights.setter
    def weights(self, val):
        pass

    def _self_normalize(self):
        s = np.sum(self.T.weights[self.T.is_leaf])
        if s > 0:
            self.T.weights[self.T.is_leaf] = self.T.weights[self.T.is_leaf] / s

    def _update_model(self):
        """Mixture of the leaf proposals (:710-736): uniform boxes for "haar", N(centre, radius * I) for "normal"."""
        leaf_nodes = np.flatnonzero(self.T.is_leaf)
        weights = self.T.weights[leaf_nodes].copy()
        models = []
        for nd in leaf_nodes:
            c, r = self.T.centers[nd], self.T.radii[nd]
            if self.kernel == "haar":
                models.append(CMultivariateUniform({"center": c.copy(), "radius": np.array([r])}))
            else:
                models.append(CMultivariateNormal({"mean": c.copy(), "sigma": np.diag(r * np.ones(self.ndims))}))
        self._self_normalize()
        self.model = CMixtureModel({"models": models, "weights": weights, "dims": self.ndims,
                                    "support": [self.space_min, self.space_max]})
        self._boxes = None
