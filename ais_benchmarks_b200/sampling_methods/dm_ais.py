"""DM-AIS / DM-PMC with the reference's interface
(reference ais_benchmarks/sampling_methods/dm_ais.py:10-89).

One iteration = N*K proposal draws (Philox kernel), ONE fused weight kernel launch
(log pi, log q over all N proposals, log w and the normaliser / ESS partials; C ABI
aisb_dm_weights_f32) and an N-draw multinomial resampling of the proposal means. The proposal
means never leave HBM; ``self.proposals`` / ``self.model`` are rebuilt as reference-style objects
only when someone reads them.
"""
import time

import numpy as np
import torch

from .. import _device
from ..distributions import CMixtureModel, CMultivariateNormal
from .base import (CMixtureISSamplingMethod, foreign_target_logp, multinomial_resample, target_device_mixture)


class CDeterministicMixtureAIS(CMixtureISSamplingMethod):
    def __init__(self, params):
        """
        :param params:
            - space_min / space_max: space boundaries; dims: number of dimensions
            - K: samples per proposal; N: number of proposals; J: iteration cap (unused, like the reference)
            - sigma: proposal VARIANCE (covariance = sigma * I, dm_ais.py:38-39)
            - sharded (optional, under torch.distributed): every rank draws and weighs its share of the K draws per
              proposal; the proposal means stay replicated (two-level multinomial resampling + an all-gather of the N
              chosen means per iteration); returned samples are this rank's rows, their weights are normalised over
              ALL ranks and iterations
            - cuda_graph (optional, default "auto"): capture one whole adapt iteration (draws, proposal packing, weight
              kernel, resampling) in a CUDA graph and replay it; "auto" does so when an iteration is launch-bound
              (at most GRAPH_MAX_BATCH draws, packed Gaussian target, not sharded); False keeps eager launches
        """
        super(CDeterministicMixtureAIS, self).__init__(params)
        self.K = params["K"]
        self.N = params["N"]
        self.J = params["J"]
        self.sigma = params["sigma"]
        self.sharded = bool(params.get("sharded", False))
        self.cuda_graph = params.get("cuda_graph", "auto")
        self._captured = None
        self._means = None
        self._mixture = None
        self._objects = None
        self._pending_records = []
        self.reset()

    # -- proposal state ----------------------------------------------------------------------------
    def reset(self):
        super(CDeterministicMixtureAIS, self).reset()
        # proposal centres ~ U(space_min, space_max) from the host RNG, as dm_ais.py:36-37
        self._pending_records = []
        centres = np.random.uniform(self.space_min, self.space_max, size=(self.N, self.ndims))
        if self.sharded:
            from .. import parallel
            centres = parallel.broadcast_array(centres)   # replicated parameters start identical on every rank
        self._n_all_ranks = 0
        self._set_means_host(centres)

    def _set_means_host(self, centres):
        self._means_host = np.asarray(centres, dtype=np.float64)
        self._means = None
        self._mixture = None
        self._objects = None

    def _set_means_dev(self, means):
        self._means = means.contiguous()
        self._means_host = None
        self._mixture = None
        self._objects = None

    def proposal_means(self):
        """float32 CUDA [N, D]."""
        if self._means is None:
            _device.require_cuda()
            self._means = torch.from_numpy(self._means_host).to(_device.device(), torch.float32).contiguous()
        return self._means

    def proposal_mixture(self):
        """Packed equal-weight isotropic mixture q = (1/N) sum_n N(mu_n, sigma I) in HBM."""
        if self._mixture is None:
            self._mixture = _device.DeviceMixture.kde(self.proposal_means(), None, self.N, float(self.sigma))
        return self._mixture

    def _build_objects(self):
        if self._objects is None:
            means = self._means_host if self._means_host is not None else _device.download(self._means)
            cov = np.diag(np.array([self.sigma] * self.ndims, dtype=np.float64))
            props = [CMultivariateNormal({"mean": m, "sigma": cov.copy()}) for m in means]
            model = CMixtureModel({"models": props, "weights": np.array([1 / len(props)] * len(props)),
                                   "dims": self.ndims, "support": [self.space_min, self.space_max]})
            self._objects = (props, model)
        return self._objects

    @property
    def proposals(self):
        return self._build_objects()[0]

    @proposals.setter
    def proposals(self, val):
        if val:
            self._set_means_host(np.array([np.asarray(p.loc, dtype=np.float64).reshape(-1) for p in val]))

    @property
    def model(self):
        return self._build_objects()[1]

    @model.setter
    def model(self, val):
        pass  # the mixture is derived from the proposal means

    # -- density / sampling of the proposal mixture ------------------------------------------------
    def log_prob(self, s):
        x, as_numpy = _device.to_device_points(s, self.ndims)
        return _device.from_device(_device.mixture_logpdf(x, self.proposal_mixture()), as_numpy)

    def prob(self, s):
        x, as_numpy = _device.to_device_points(s, self.ndims)
        return _device.exp_from_device(_device.mixture_logpdf(x, self.proposal_mixture()), as_numpy)

    def sample(self, n_samples):
        self._num_q_samples += n_samples
        means = self.proposal_means()
        pick = torch.randint(0, self.N, (int(n_samples),), device=means.device)
        x = _device.sample_iso(means[pick].contiguous(), 1, float(self.sigma))
        return x if self.device_outputs else _device.download(x)

    # -- the adapt loop ----------------------------------------------------------------------------
    def weigh(self, new_samples, target_d, hint=None):
        """log w = log pi(x) - log q(x) for a batch + its normaliser/ESS partials (dm_ais.py:72-76).
        hint: int32 CUDA [n], the proposal each sample was drawn from; in 9..32 dimensions it routes the proposal
        density through the tensor-core kernel (same values, see csrc/tc.cu)."""
        target = target_device_mixture(target_d)
        logp_in = None if target is not None else foreign_target_logp(target_d, new_samples)
        logw, _, _, partials = _device.dm_weights(new_samples, target, logp_in, self.proposal_mixture(), hint=hint)
        if hint is not None:
            # the hinted path may run the tensor-core kernel, which reports a barrier time-out only through the record
            # (n_valid = -1) and leaves log q unwritten: remember the record, importance_sample() checks it where it
            # reads results back anyway (stream-ordered, one small D2H per call)
            self._pending_records.append(partials)
        return logw, partials

    def _check_pending_records(self):
        """Raise AisbError if any weighting since the last check aborted (see weigh)."""
        if self._pending_records:
            recs, self._pending_records = torch.stack(self._pending_records), []
            _device.check_partials(recs)

    def _own_proposal(self, per=None):
        """Proposal index of every row of a proposal-major batch of N*per draws."""
        per = int(self.K if per is None else per)
        if getattr(self, "_hint", None) is None or self._hint.shape[0] != self.N * per:
            self._hint = (torch.arange(self.N * per, device=_device.device()) // max(per, 1)).to(torch.int32)
        return self._hint

    def _draws_per_proposal(self):
        """K, or this rank's share of the K draws per proposal when sharded."""
        if not self.sharded:
            return int(self.K)
        from .. import parallel
        a, b = parallel.shard_rows(int(self.K))
        return b - a

    def resample(self, samples, logw):
        """DM-PMC update (dm_ais.py:47-56): each proposal mean <- a sample drawn ~ batch-normalised weights (over the
        batches of all ranks when sharded)."""
        if self.sharded:
            from .. import parallel
            self._set_means_dev(parallel.resample_global(samples, logw, self.N))
            return
        idx = multinomial_resample(logw, self.N)
        self._set_means_dev(samples[idx])

    def _global_partials(self):
        """Weight record of the accumulated log weights, merged over the ranks when sharded."""
        rec = self._weight_partials()
        if self.sharded:
            from .. import parallel
            rec = parallel.global_weight_partials(rec)
        return rec

    def get_approx_NESS(self):
        if not self.sharded:
            return super(CDeterministicMixtureAIS, self).get_approx_NESS()
        rec = self._global_partials()
        return _device.weight_ess(rec) / rec[3]

    # -- one adapt iteration as a replayed CUDA graph ------------------------------------------------
    GRAPH_MAX_BATCH = 1 << 18   # draws per iteration up to which an iteration is launch-bound on a B200

    def _graph_wanted(self, target_d, per):
        if self.cuda_graph is False or self.cuda_graph == "off" or self.sharded or per < 1:
            return False
        if target_device_mixture(target_d) is None:
            return False          # a foreign target runs host code inside the iteration
        return self.cuda_graph is True or self.N * per <= self.GRAPH_MAX_BATCH

    def _adapt_in_graph(self, mbuf, x, logw, target):
        """The adaptation step of the captured iteration, in place on the proposal means `mbuf` (DM-PMC resampling,
        dm_ais.py:47-56). Overridden by LAIS."""
        mbuf.copy_(x[multinomial_resample(logw, self.N)])

    def _iteration_ops(self, mbuf, target, per, ws):
        """Every device operation of one iteration, on buffers the caller owns (capturable: no host synchronisation,
        Philox position in device memory). The proposal density runs on the CUDA cores here -- the batches this path is
        meant for are far too small for the tensor-core kernel to matter."""
        x = _device.sample_iso_ctr(mbuf, per, float(self.sigma))
        mix = _device.DeviceMixture.kde(mbuf, None, self.N, float(self.sigma))
        logw, _, _, part = _device.dm_weights(x, target, None, mix, ws=ws)
        self._adapt_in_graph(mbuf, x, logw, target)
        return x, logw, part, mix

    def _prepare_graph(self, dev):
        """Hook: device constants an iteration needs, created eagerly BEFORE the capture (host-to-device copies from
        pageable memory are not capturable)."""

    def _captured_iteration(self, target_d, per):
        """Build (once per target / shape) and return the captured iteration."""
        target = target_device_mixture(target_d)
        key = (target.rows.data_ptr(), target.K, int(self.N), int(per), float(self.sigma), int(self.ndims))
        cap = self._captured
        if cap is None or cap["key"] != key:
            dev = _device.device()
            self._prepare_graph(dev)
            means = torch.empty((self.N, self.ndims), dtype=torch.float32, device=dev)
            ws = torch.empty(_device.workspace_bytes_for(self.N * per, max(self.N, target.K), self.ndims) + 256,
                             dtype=torch.uint8, device=dev)
            # warm-up and capture must not consume random numbers: a sampler that captures now and one that captured
            # earlier draw the same streams after the same seed
            saved_philox = _device.rng_state().clone()
            saved_torch = torch.cuda.get_rng_state(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up on a scratch copy: loads the kernels, adapts nothing
                scratch = self.proposal_means().clone()
                for _ in range(2):
                    self._iteration_ops(scratch, target, per, ws)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                x, logw, part, mix = self._iteration_ops(means, target, per, ws)
            _device.rng_state().copy_(saved_philox)
            torch.cuda.set_rng_state(saved_torch, dev)
            cap = {"key": key, "graph": graph, "means": means, "x": x, "logw": logw, "part": part,
                   "keep": (mix, ws, target)}
            self._captured = cap
        if self._means is not cap["means"]:          # reset() / a caller replaced the proposal means
            cap["means"].copy_(self.proposal_means())
            self._set_means_dev(cap["means"])
        return cap

    def _graph_iteration(self, target_d, per):
        try:
            cap = self._captured_iteration(target_d, per)
        except RuntimeError as e:
            # capture refused (e.g. another thread is using the device): this sampler stays on eager launches
            import warnings
            warnings.warn("CUDA-graph capture of the adapt iteration failed (%s); using eager launches" % e)
            self.cuda_graph = False
            self._captured = None
            torch.cuda.synchronize()
            return False
        cap["graph"].replay()
        # the means changed in place: everything derived from them is stale
        self._mixture = None
        self._objects = None
        self._num_q_samples += self.N * per
        self._num_pi_evals += self.N * per
        # the graph's outputs are overwritten by the next replay: the accumulated state gets copies
        self._append(cap["x"].clone(), cap["logw"].clone())
        self._n_all_ranks += self.N * int(self.K)
        return True

    def importance_sample(self, target_d, n_samples, timeout=60):
        elapsed_time = 0
        t_ini = time.time()
        per = self._draws_per_proposal()
        # n_samples is the total wanted over all ranks; every iteration adds N*K of them
        # (sharded: the loop is driven by the sample count alone -- a per-rank clock must not desynchronise the collectives)
        while n_samples > (self._n_all_ranks if self.sharded else self._n()) and (self.sharded or elapsed_time < timeout):
            if self._graph_wanted(target_d, per) and self._graph_iteration(target_d, per):
                elapsed_time = time.time() - t_ini
                continue
            # K samples from each proposal, proposal-major like dm_ais.py:64-69
            new_samples = _device.sample_iso(self.proposal_means(), per, float(self.sigma))
            self._num_q_samples += self.N * per
            new_logw, _ = self.weigh(new_samples, target_d, hint=self._own_proposal(per))
            self._num_pi_evals += self.N * per
            self.resample(new_samples, new_logw)
            self._append(new_samples, new_logw)
            self._n_all_ranks += self.N * int(self.K)
            elapsed_time = time.time() - t_ini
        self._check_pending_records()
        # weights / weights.sum() over ALL iterations (dm_ais.py:86) -- and all ranks
        norm_w = _device.normalize_weights(self._logw_dev, self._global_partials())
        return self._out(self._samples_dev), (norm_w if self.device_outputs else _device.download(norm_w))

    def _update_model(self):
        pass
