"""Host-side behaviour of the reference-API classes that needs no GPU: constructor contracts, parameter
sanitising, shape checks, counters / reset, and the loud failure when CUDA is absent."""
import numpy as np
import pytest
import torch

from ais_benchmarks_b200 import AisbError
from ais_benchmarks_b200._lib import COV_DIAG, COV_FULL, COV_ISO_SHARED
from ais_benchmarks_b200.distributions import (CGaussianMixtureModel, CMixtureModel, CMultivariateNormal,
                                               CMultivariateUniform)
from ais_benchmarks_b200.distributions.mixture.CMixtureModel import gaussian_pack_args, sanitize_weights
from ais_benchmarks_b200.metrics import CKLDivergence, CMetric
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS, CMixturePMC

no_gpu = pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-CUDA failure mode")


def test_mvn_constructor_contract():
    d = CMultivariateNormal({"mean": np.array([0.0, 1.0]), "sigma": np.array([[2.0, 0.3], [0.3, 1.0]])})
    assert d.dims == 2 and d.type == "normal" and d.family == "exponential"
    assert np.allclose(d.inv_cov, np.linalg.inv(d.scale)) and d.det == pytest.approx(1.91)
    assert d.term1 == pytest.approx(-np.log(2 * np.pi)) and d.term2 == pytest.approx(-0.5 * np.log(1.91))
    assert np.allclose(d.support(), [[-6 * np.sqrt(2), 1 - 6], [6 * np.sqrt(2), 1 + 6]])   # CMultivariateNormal.py:37-40
    v0 = d._version
    d.set_moments(np.zeros(2), np.eye(2))
    assert d._version > v0 and d.is_diagonal()           # every parameter change invalidates packed copies
    v1 = d._version
    d.loc = np.ones(2)                                         # the public attribute setters invalidate them too
    assert d._version > v1 and d._packed is None
    with pytest.raises(AssertionError):
        d.set_moments(np.zeros(2), np.zeros((2, 2)))           # det > 0 (CMultivariateNormal.py:55-56)
    with pytest.raises(AssertionError):
        d.set_moments(np.zeros(2), np.eye(3))
    with pytest.raises(AssertionError):
        CMultivariateNormal({"mean": np.zeros(2)})             # missing "sigma"
    with pytest.raises(NotImplementedError):
        d.marginal(0)
    with pytest.raises(ValueError):
        d._check_shape(np.zeros((4, 3)))
    assert d._check_shape(np.zeros(2)).shape == (1, 2)


def test_weight_sanitising_matches_reference_rules(capsys):
    w = np.array([0.3, 0.0, np.nan, 0.5, -1.0, 0.2])
    out = sanitize_weights(w)
    assert np.allclose(out, [0.3, 0, 0, 0.5, 0, 0.2])           # sums to 1 already
    assert np.array_equal(w[[1, 2, 4]], [0, 0, 0])               # the caller's array is modified in place (CMixtureModel.py:33)
    assert "3 NaN, negative or zero weights out of 6" in capsys.readouterr().err
    z = sanitize_weights(np.zeros(3))
    assert np.array_equal(z, np.zeros(3))
    assert "All weights are set to zero" in capsys.readouterr().err
    assert np.allclose(sanitize_weights(np.array([2.0, 6.0])), [0.25, 0.75])


def test_mixture_and_gmm_constructors():
    g = CGaussianMixtureModel({"means": [[1.5], [-4.2], [0.01], [7.5]], "sigmas": [[0.4], [0.05], [0.003], [0.003]],
                               "weights": [0.4, 0.4, 0.1, 0.1], "support": [[-10], [10]]})
    assert g.dims == 1 and len(g.models) == 4 and g.weights == [0.4, 0.4, 0.1, 0.1]     # raw weights are kept
    assert np.allclose(g.gmm.weights, [0.4, 0.4, 0.1, 0.1]) and g.support().shape == (2, 1)
    assert g.models[1].scale.shape == (1, 1) and g.models[1].scale[0, 0] == 0.05        # variances on the diagonal
    d = g.to_dict(name="t")
    assert d["type"] == "distributions.CGaussianMixtureModel" and d["params"]["means"][3] == [7.5]
    g2 = CGaussianMixtureModel({"means": np.zeros((3, 2)), "sigmas": np.ones((3, 2)), "support": [[-1, -1], [1, 1]]})
    assert np.allclose(g2.weights, 1 / 3)
    with pytest.raises(AssertionError):
        CMixtureModel({"models": g.models, "weights": [1.0], "dims": 1, "support": [[-1], [1]]})
    with pytest.raises(AssertionError):
        CMixtureModel({"models": g.models, "weights": [1, 1, 1, 1], "support": [[-1], [1]]})   # "dims" is required
    # layout choice of the packer
    iso = [CMultivariateNormal({"mean": np.zeros(3) + i, "sigma": np.eye(3) * 0.2}) for i in range(4)]
    assert gaussian_pack_args(iso)[2] == COV_ISO_SHARED
    iso[1].set_moments(np.ones(3), np.diag([0.2, 0.3, 0.2]))
    assert gaussian_pack_args(iso)[2] == COV_DIAG
    iso[2].set_moments(np.ones(3), np.array([[1, .1, 0], [.1, 1, 0], [0, 0, 1.0]]))
    assert gaussian_pack_args(iso)[2] == COV_FULL
    m = CMixtureModel({"models": iso, "weights": np.ones(4), "dims": 3, "support": [[-9] * 3, [9] * 3]})
    k1 = m._key()
    iso[0].set_moments(np.zeros(3), np.eye(3))
    assert m._key() != k1                                          # component edits invalidate the packed cache


def test_uniform_constructor():
    u = CMultivariateUniform({"center": np.array([0.0, 0.1]), "radius": np.array([0.5, 0.2])})
    assert u.prob_val == pytest.approx(1 / (1.0 * 0.4)) and np.allclose(u.support(), [[-0.5, -0.1], [0.5, 0.3]])
    u1 = CMultivariateUniform({"center": np.zeros(3), "radius": np.array([0.5])})
    lo, hi, lpv = u1.box()
    assert np.allclose(lo, -0.5) and np.allclose(hi, 0.5) and lpv == pytest.approx(0.0)
    with pytest.raises(ValueError):
        CMultivariateUniform({"center": np.zeros(3), "radius": np.array([0.5, 0.5])})
    assert u.integral(np.array([-1.0, -1.0]), np.array([0.0, 1.0])) == pytest.approx(0.5)


def test_sampler_state_and_counters():
    np.random.seed(3)
    base = {"space_min": np.array([-1.0, -2.0]), "space_max": np.array([1.0, 2.0]), "dims": 2}
    dm = CDeterministicMixtureAIS(dict(base, K=3, N=4, J=10, sigma=0.1, name="dm"))
    assert dm.name == "dm" and dm.get_stats() == {"proposal_samples": 0, "proposal_evals": 0, "target_evals": 0,
                                                   "n_samples": 0}
    means = np.array([p.loc for p in dm.proposals])
    assert means.shape == (4, 2) and np.all(means >= base["space_min"]) and np.all(means <= base["space_max"])
    assert np.allclose(dm.proposals[0].scale, np.eye(2) * 0.1) and np.allclose(dm.model.weights, 0.25)
    dm.name, dm.debug = "x", True
    assert dm.name == "x" and dm.debug is True
    assert len(dm.samples) == 0 and len(dm.weights) == 0
    with pytest.raises(AssertionError):
        dm.get_acceptance_rate()
    dm.reset()
    assert not np.allclose(np.array([p.loc for p in dm.proposals]), means)      # new proposal centres after reset
    with pytest.raises(AssertionError):
        CDeterministicMixtureAIS(dict(base, dims=3, K=3, N=4, J=10, sigma=0.1))
    pmc = CMixturePMC(dict(base, K=5, N=3, J=10, sigma=0.01))
    assert np.allclose(pmc.wproposals, 1 / 3) and len(pmc.proposals) == 3 and not pmc.paper_covariance
    la = CLayeredAIS(dict(base, K=2, N=3, J=10, L=4, sigma=0.1, mh_sigma=0.5))
    assert la.L == 4 and la.mhsigma == 0.5 and len(la.proposals) == 3


def test_metric_contract():
    m = CKLDivergence()
    assert (m.name, m.type, m.is_symmetric, m.disjoint_support, m.value) == ("KL", "divergence", False, False, 0)
    m.pre(); m.post(); m.reset()
    with pytest.raises(AssertionError):
        m.compute(p=object(), q=object(), nsamples=10)
    with pytest.raises(NotImplementedError):
        CMetric().compute()
    res = CKLDivergence.compute_from_probs(np.array([0.2, 0.3, 0.5, 0.0]), np.array([0.25, 0.25, 0.5, 0.1]))
    p, q = np.array([0.2, 0.3, 0.5]), np.array([0.25, 0.25, 0.5])
    assert res.sum() == pytest.approx(np.sum(p * np.log(p / q)))


@no_gpu
def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point raises instead of silently running elsewhere."""
    g = CGaussianMixtureModel({"means": [[0.0]], "sigmas": [[1.0]], "support": [[-5], [5]]})
    for call in (lambda: g.log_prob(np.zeros((2, 1))), lambda: g.prob(np.zeros((2, 1))), lambda: g.sample(3),
                 lambda: CMultivariateUniform({"center": np.zeros(1), "radius": np.ones(1)}).log_prob(np.zeros((1, 1))),
                 lambda: CKLDivergence().compute(p=g, q=g, nsamples=10)):
        with pytest.raises(AisbError, match="no CPU fallback"):
            call()
    np.random.seed(0)
    dm = CDeterministicMixtureAIS({"space_min": np.array([-1.0]), "space_max": np.array([1.0]), "dims": 1, "K": 2, "N": 2,
                                   "J": 1, "sigma": 0.1})
    with pytest.raises(AisbError):
        dm.importance_sample(g, 4)


class ReplayDraws:
    """Feeds TP-AIS the prototype draws captured from the reference run (tests/golden/make_golden.py:make_tpais)."""

    def __init__(self, rows):
        self.rows, self.pos = rows, 0

    def _take(self, n, dims):
        out = self.rows[self.pos:self.pos + n]
        assert out.shape == (n, dims), "replay exhausted or out of step with the reference"
        self.pos += n
        return out

    uniform = _take
    normal = _take


class OracleTarget:
    """float64 host target: isolates the tree logic from float32 kernel rounding in the CPU test."""

    def __init__(self, means, sigmas, weights):
        self.args = (np.asarray(means, float), np.asarray(sigmas, float), np.asarray(weights, float))

    def prob(self, x):
        from oracle import ais_oracle as O
        return O.gmm_prob(x, *self.args)


def run_tpais_replay(case, target):
    from conftest import load_golden
    from ais_benchmarks_b200.sampling_methods import CTreePyramidSampling
    g = load_golden("tpais")
    ess_target, n_min, par, n_samples = g["cfg%d" % case]
    s = CTreePyramidSampling({"space_min": np.array([-1.0, -1.0]), "space_max": np.array([1.0, 1.0]), "dims": 2,
                              "method": "simple", "resampling": "leaf", "kernel": str(g["kernel%d" % case]),
                              "ess_target": float(ess_target), "n_min": int(n_min), "parallel_samples": int(par)})
    s.draws = ReplayDraws(g["draws%d" % case])
    samples, weights = s.importance_sample(target, int(n_samples), timeout=600)
    return g, s, samples, weights


@pytest.mark.parametrize("case", range(3))
def test_tpais_tree_logic_replays_reference_run(case):
    """Same draws + same (float64) target -> the same tree, samples, weights, NESS and counters as the reference
    (sampling_methods/tree_pyramid.py:421-457, 500-533, 646-708), including the `_get_samples` offset quirk."""
    g, s, samples, weights = run_tpais_replay(case, OracleTarget([[0.0, 0.0], [-0.3, 0.5]], [[0.01, 0.1], [0.01, 0.03]],
                                                                   [0.5, 0.5]))
    assert s.draws.pos == len(g["draws%d" % case])                      # consumed exactly the reference's draws
    assert np.allclose(s.T.centers, g["centers%d" % case], atol=1e-15)
    assert np.allclose(s.T.radii, g["radii%d" % case], atol=1e-15)
    assert np.array_equal(s.T.leaves_idx, g["leaves_idx%d" % case])
    assert np.allclose(s.T.weights, g["tweights%d" % case], rtol=1e-9, atol=1e-300)
    assert samples.shape == g["samples%d" % case].shape
    assert np.allclose(samples, g["samples%d" % case], atol=1e-12)
    assert np.allclose(weights, g["weights%d" % case], rtol=1e-9)
    assert s.get_NESS() == pytest.approx(float(g["ness%d" % case]), rel=1e-9)
    assert [s.num_proposal_samples, s.num_proposal_evals, s.num_target_evals] == list(g["counters%d" % case])
    assert s.get_acceptance_rate() == pytest.approx(float(g["acc%d" % case]))
    assert np.allclose(s.model.weights, g["tweights%d" % case][g["leaves_idx%d" % case]], rtol=1e-9)


def test_tpais_rejects_dead_reference_methods():
    from ais_benchmarks_b200.sampling_methods import CTreePyramidSampling
    base = {"space_min": np.array([-1.0]), "space_max": np.array([1.0]), "dims": 1, "resampling": "leaf",
            "kernel": "haar", "ess_target": 0.9, "n_min": 5, "parallel_samples": 4}
    with pytest.raises(NotImplementedError):
        CTreePyramidSampling(dict(base, method="dm"))
    with pytest.raises(AssertionError):
        CTreePyramidSampling(dict(base, method="bogus"))
    with pytest.raises(AssertionError):
        CTreePyramidSampling(dict(base, method="simple", kernel="triangle"))


def test_benchmark_loaders_and_results_round_trip(tmp_path):
    """benchmark/CBenchmark.py: the reference's YAML schema (CBenchmark.py:58-137), its results.txt row format (:318-326)
    and the skip-with-a-note rule for methods that are not on the accelerated path."""
    import os
    from ais_benchmarks_b200.benchmark import CBenchmark
    from ais_benchmarks_b200.benchmark.CBenchmark import _literal, time_to_hms
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ais_benchmarks_b200", "benchmark")
    b = CBenchmark()
    b.load_config(os.path.join(here, "def_config.yaml"))
    assert b.metrics == ["KLD", "JSD", "EVMSE", "T"] and b.n_experiments == 1 and b.rseed == 0
    b.load_benchmark(os.path.join(here, "def_benchmark.yaml"))
    assert [t.name for t in b.targets] == ["gmm", "gmm", "normal"] and b.ndims == [1, 2, 1]
    assert b.nsamples == [500, 1000, 1000] and b.eval_sampl == [2000] * 3 and b.batch_sizes == [2] * 3
    assert isinstance(b.targets[0], CGaussianMixtureModel) and isinstance(b.targets[2], CMultivariateNormal)
    assert _literal("'simple'") == "simple" and _literal(0.9) == 0.9 and _literal("[1, 2]") == [1, 2]
    assert _literal("np.zeros(3)") == "np.zeros(3)"  # not a literal: left alone, never evaluated
    assert time_to_hms(3723.5) == (1, 2, 3.5)
    bad = tmp_path / "bad.yaml"
    bad.write_text("targets:\n  - {name: x, type: Banana2D, params: {}, batch_size: 1, nsamples: 1, nsamples_eval: 1}\n")
    with pytest.raises(ValueError):
        b.load_benchmark(str(bad))

    class Method:
        name, num_proposal_samples, num_proposal_evals, num_target_evals = "dm_ais", 150, 750, 150
        get_NESS = staticmethod(lambda: 0.4321)
        get_acceptance_rate = staticmethod(lambda: 1.0)

    class Target:
        dims, name = 2, "gmm"

    out = tmp_path / "results.txt"
    out.write_text(CBenchmark.results_header(["KLD", "JSD", "T"]))
    CBenchmark.write_results({"KL": 1.25, "JSD": 0.5, "time": 0.001}, Method, Target, 150, str(out))
    CBenchmark.write_results({"KL": 0.75, "JSD": 0.25, "time": 0.002}, Method, Target, 300, str(out))
    lines = out.read_text().splitlines()
    assert lines[0] == "dims output_samples KLD JSD T NESS method target_d accept_rate proposal_samples proposal_evals target_evals"
    assert lines[1] == "02 0150   1.25000   0.50000   0.00100    0.432 dm_ais gmm 1.000 150 750 150"
    table = CBenchmark.read_results(str(out))
    assert table["output_samples"] == [150.0, 300.0] and table["KLD"] == [1.25, 0.75] and table["method"] == ["dm_ais"] * 2


@no_gpu
def test_benchmark_methods_and_no_cpu_fallback(tmp_path):
    import os
    from ais_benchmarks_b200.benchmark import CBenchmark
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ais_benchmarks_b200", "benchmark")
    y = tmp_path / "m.yaml"
    y.write_text("methods:\n  - {name: rej, type: CRejectionSampling, debug: false, params: {scaling: 1}}\n")
    b = CBenchmark()
    b.load_methods(str(y), np.array([-1.0]), np.array([1.0]), 1)
    assert b.methods == [] and b.skipped_methods == ["rej"]
    b.load_methods(os.path.join(here, "def_methods.yaml"), np.array([-1.0]), np.array([1.0]), 1)
    assert [m.name for m in b.methods] == ["m_pmc", "TP_AISr_ESS90", "lais", "dm_ais"] and b.skipped_methods == []
    assert b.methods[1].method == "simple" and b.methods[3].K == 10
    b.load_benchmark(os.path.join(here, "def_benchmark.yaml"))
    with pytest.raises(AisbError):  # no CPU fallback: the evaluation loop fails loudly without CUDA
        CBenchmark.evaluate_method(1, b.targets[2], b.methods[3], max_samples=50, sampling_eval_samples=100,
                                   metrics=["KLD"], n_reps=1, batch_size=2)


# ---------------------------------------------------------------------------------------------------
# Host-side logic added around the kernels: Z-order permutation, sync-free resampling, spatial sharding
# (device-agnostic torch code paths, exercised here on CPU tensors)
# ---------------------------------------------------------------------------------------------------
def test_morton_order_reference_path_cpu():
    import torch
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(0)
    for D in (1, 2, 3, 8, 16, 33):
        x = torch.from_numpy(rs.uniform(-2, 5, size=(4000, D)).astype(np.float32))
        x[100:200] = x[0]
        perm = _device.morton_order(x)                       # CPU tensors take the torch reference path
        assert sorted(perm.tolist()) == list(range(4000))
        assert _device.morton_bits(D) * D <= max(32, D)      # D > 32: one bit per coordinate (keys only used for D <= 16)
    # locality: consecutive rows along the curve are much closer than consecutive rows in random order
    x = torch.from_numpy(rs.uniform(0, 1, size=(20000, 2)).astype(np.float32))
    p = _device.morton_order(x)
    step_sorted = (x[p][1:] - x[p][:-1]).norm(dim=1).mean()
    step_random = (x[1:] - x[:-1]).norm(dim=1).mean()
    assert step_sorted < 0.1 * step_random


def test_multinomial_resample_is_sync_free_and_handles_degenerate_weights():
    import torch
    from ais_benchmarks_b200.sampling_methods.base import multinomial_resample
    torch.manual_seed(0)
    logw = torch.log(torch.tensor([0.1, 0.0, 0.6, 0.3]))
    logw[1] = float("nan")                                   # invalid weights count as zero
    idx = multinomial_resample(logw, 40000)
    freq = torch.bincount(idx, minlength=4).double() / 40000
    assert freq[1] == 0 and torch.allclose(freq, torch.tensor([0.1, 0.0, 0.6, 0.3], dtype=torch.float64), atol=0.01)
    # +inf weights are invalid too; a single valid entry takes every draw
    idx = multinomial_resample(torch.tensor([float("inf"), -3.0, float("-inf")]), 100)
    assert bool((idx == 1).all())
    # nothing valid: uniform over all rows (the reference would divide by zero)
    idx = multinomial_resample(torch.full((5,), float("-inf")), 50000)
    freq = torch.bincount(idx, minlength=5).double() / 50000
    assert torch.allclose(freq, torch.full((5,), 0.2, dtype=torch.float64), atol=0.01)
    assert multinomial_resample(torch.zeros(3), 0).shape == (0,)


def test_spatial_shard_partitions_the_point_set_cpu(monkeypatch):
    import torch
    from ais_benchmarks_b200 import parallel
    monkeypatch.setattr(parallel, "SPATIAL_BLOCK", 16)
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.normal(size=(1001, 3)).astype(np.float32))
    seen = torch.zeros(1001, dtype=torch.int64)
    sizes = []
    for r in range(4):
        rows, idx = parallel.spatial_shard(x, r, 4)
        assert torch.equal(rows, x[idx])
        seen[idx] += 1
        sizes.append(len(idx))
    assert bool((seen == 1).all()) and max(sizes) - min(sizes) <= 16     # block-cyclic: balanced to within one block
    # every block a rank receives is a run of curve neighbours: its bounding box is much smaller than the whole set's,
    # while the rank as a whole sees every region of space (that is what balances the refinement load)
    whole = (x.max(0).values - x.min(0).values).prod()
    rows, _ = parallel.spatial_shard(x, 1, 4)
    block = rows[:16]
    assert (block.max(0).values - block.min(0).values).prod() < 0.1 * whole
    assert (rows.max(0).values - rows.min(0).values).prod() > 0.5 * whole
    one, idx1 = parallel.spatial_shard(x, 0, 1)                          # a single rank: the whole curve, in order
    assert len(idx1) == 1001 and torch.equal(one, x[idx1])


def test_peer_exchange_is_off_without_nccl():
    from ais_benchmarks_b200 import parallel
    parallel.PeerRecordExchange._instance = None
    parallel.PeerRecordExchange._disabled = False
    assert parallel.PeerRecordExchange.get() is None        # single process, no process group: NCCL path / no exchange
    parallel.PeerRecordExchange._disabled = False


def test_device_path_follows_the_class_that_defines_log_prob():
    """metrics.divergences keeps an object's log_prob on the device iff the method it will call is this package's:
    a user's subclass (the reference's extension mechanism: samplers and distributions are subclassed by callers) must
    not be demoted to the numpy round trip, an override or a foreign class must."""
    from ais_benchmarks_b200.metrics.divergences import _is_ours
    from ais_benchmarks_b200.sampling_methods.base import CMixtureSamplingMethod
    from ais_benchmarks_b200.distributions import CMultivariateNormal

    class UserSampler(CMixtureSamplingMethod):
        def importance_sample(self, target_d, n_samples, timeout):
            raise NotImplementedError

    class UserNormal(CMultivariateNormal):
        def log_prob(self, x):
            return np.zeros(len(x))

    class Foreign:
        def log_prob(self, x):
            return np.zeros(len(x))

    params = {"space_min": np.zeros(2), "space_max": np.ones(2), "dims": 2, "n_samples_kde": 10}
    assert _is_ours(UserSampler(params))
    assert _is_ours(CMultivariateNormal({"mean": np.zeros(2), "sigma": np.eye(2)}))
    assert not _is_ours(UserNormal({"mean": np.zeros(2), "sigma": np.eye(2)}))
    assert not _is_ours(Foreign())
    assert not _is_ours(object())


def test_upload_thread_count_policy(monkeypatch):
    """Host threads of the staged uploader / downloader: explicit override clamped to the library's 1..8 lanes, else
    the cores one local rank can claim (torchrun's LOCAL_WORLD_SIZE), at most 8, never 0."""
    import os
    from ais_benchmarks_b200 import _device
    monkeypatch.delenv("AISB_UPLOAD_THREADS", raising=False)
    monkeypatch.setattr(os, "cpu_count", lambda: 16)
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    assert _device._upload_threads() == 8
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert _device._upload_threads() == 2
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "64")
    assert _device._upload_threads() == 1
    monkeypatch.setattr(os, "cpu_count", lambda: None)
    assert _device._upload_threads() == 1
    monkeypatch.setenv("AISB_UPLOAD_THREADS", "5")
    assert _device._upload_threads() == 5
    monkeypatch.setenv("AISB_UPLOAD_THREADS", "99")
    assert _device._upload_threads() == 8
    monkeypatch.setenv("AISB_UPLOAD_THREADS", "0")
    assert _device._upload_threads() == 1
