"""Parity of the CUDA path (through the Python API -> C ABI -> sm_100a kernels) against the committed
reference fixtures and the oracle. Tolerances (BASELINE.json north_star): log densities 1e-5,
KLD and normalised weights 1e-4, both as |got - ref| / max(1, |ref|)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ais_oracle as O

pytestmark = pytest.mark.gpu

LOGP_TOL = 1e-5
W_TOL = 1e-4
KLD_TOL = 1e-4


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ais_benchmarks_b200 import _lib
    _lib.load()


def gmm(means, sigmas, weights, support=None):
    from ais_benchmarks_b200.distributions import CGaussianMixtureModel
    d = means.shape[1]
    sup = support if support is not None else [np.full(d, -10.0), np.full(d, 10.0)]
    return CGaussianMixtureModel({"means": means, "sigmas": sigmas, "weights": weights, "support": sup})


# ---------------------------------------------------------------------------------------------------
def test_mvn_golden_and_shapes():
    from ais_benchmarks_b200.distributions import CMultivariateNormal
    g = load_golden("mvn")
    for i in range(int(g["n"])):
        dist = CMultivariateNormal({"mean": g["mean%d" % i].copy(), "sigma": g["cov%d" % i].copy()})
        x = g["x%d" % i]
        lp = dist.log_prob(x)
        assert lp.shape == (len(x),) and lp.dtype == np.float64
        assert rel_err(lp, g["logp%d" % i]) < LOGP_TOL
        assert np.allclose(dist.prob(x), g["prob%d" % i], rtol=1e-4, atol=0)
        assert np.allclose(dist.support(), g["support%d" % i])
        # unbatched point -> (1,) result (distributions.py:137-138)
        one = dist.log_prob(x[0])
        assert one.shape == (1,) and abs(one[0] - g["logp%d" % i][0]) < LOGP_TOL * max(1, abs(g["logp%d" % i][0]))
        # torch in -> torch out, stays on the device
        lt = dist.log_prob(torch.from_numpy(x).cuda())
        assert isinstance(lt, torch.Tensor) and lt.is_cuda
        assert rel_err(lt.cpu().numpy(), g["logp%d" % i]) < LOGP_TOL
    with pytest.raises(ValueError):
        dist.log_prob(np.zeros((3, dist.dims + 1)))
    with pytest.raises(ValueError):
        dist.log_prob(np.zeros((2, 2, dist.dims)))
    with pytest.raises(AssertionError):
        CMultivariateNormal({"mean": np.zeros(2), "sigma": np.array([[1.0, 2.0], [2.0, 1.0]])})
    with pytest.raises(AssertionError):
        CMultivariateNormal({"sigma": np.eye(2)})


@pytest.mark.parametrize("case", range(6))
def test_gmm_golden(case):
    g = load_golden("gmm")
    dist = gmm(g["means%d" % case], g["sigmas%d" % case], g["weights%d" % case], g["support%d" % case])
    x = g["x%d" % case]
    assert rel_err(dist.log_prob(x), g["logp%d" % case]) < LOGP_TOL
    ref_p = g["prob%d" % case]
    got_p = dist.prob(x)
    nz = ref_p > 1e-300
    assert np.allclose(got_p[nz], ref_p[nz], rtol=2e-4, atol=0)
    assert np.all(got_p[~nz] < 1e-290)


def test_gmm_from_yaml_lists():
    """def_benchmark.yaml passes plain lists (benchmark/CBenchmark.py:129-131)."""
    from ais_benchmarks_b200.distributions import CGaussianMixtureModel
    g = load_golden("gmm")
    dist = CGaussianMixtureModel({"means": [[1.5], [-4.2], [0.01], [7.5]], "sigmas": [[0.4], [0.05], [0.003], [0.003]],
                                  "weights": [0.4, 0.4, 0.1, 0.1], "support": [[-10], [10]]})
    assert dist.dims == 1 and dist.support().shape == (2, 1)
    assert rel_err(dist.log_prob(g["x0"]), g["logp0"]) < LOGP_TOL


def test_mixture_bad_weights_full_cov():
    from ais_benchmarks_b200.distributions import CMixtureModel, CMultivariateNormal
    g = load_golden("mixture")
    models = [CMultivariateNormal({"mean": m.copy(), "sigma": c.copy()}) for m, c in zip(g["means"], g["covs"])]
    mix = CMixtureModel({"models": models, "weights": g["weights"].copy(), "dims": 3, "support": [[-2] * 3, [2] * 3]})
    assert np.allclose(mix.weights, g["norm_weights"])
    assert rel_err(mix.log_prob(g["x"]), g["logp"]) < LOGP_TOL
    assert np.allclose(mix.prob(g["x"]), g["prob"], rtol=2e-4)
    assert rel_err(mix.log_prob(g["x"][0]), g["logp_single"]) < LOGP_TOL
    # parameters changed through a component object are picked up (dm_ais.py:56 style set_moments)
    models[0].set_moments(g["means"][0] + 0.5, g["covs"][0])
    means2 = g["means"].copy()
    means2[0] += 0.5
    assert rel_err(mix.log_prob(g["x"]), O.mixture_log_prob(g["x"], means2, g["covs"], g["weights"])) < LOGP_TOL
    # all weights zero -> -inf everywhere (logsumexp with b = 0)
    mix.set_weights(np.zeros(6))
    assert np.all(np.isneginf(mix.log_prob(g["x"][:5])))


@pytest.mark.parametrize("D", [1, 2, 3, 4, 5, 8, 12, 16, 32, 40, 64])
@pytest.mark.parametrize("kind", ["iso", "diag", "full"])
def test_every_kernel_instantiation_vs_oracle(D, kind):
    """One small case per (covariance kind, padded D) template instantiation, incl. non power-of-two D."""
    from ais_benchmarks_b200 import _device, _lib
    if kind == "full" and D > 32:
        with pytest.raises(_lib.AisbError):
            _device.DeviceMixture.from_params(np.zeros((1, D)), np.eye(D)[None], None, _lib.COV_FULL)
        return
    rs = np.random.RandomState(D * 7 + len(kind))
    K, n = 37, 700
    means = rs.uniform(-1, 1, size=(K, D))
    w = rs.uniform(0.1, 1, size=K)
    x = rs.uniform(-1.2, 1.2, size=(n, D))
    x[:K] = means + 0.05 * rs.normal(size=(K, D))
    if kind == "iso":
        v = 0.07
        mix = _device.DeviceMixture.from_params(means, np.array([v]), w, _lib.COV_ISO_SHARED)
        ref = O.iso_mixture_log_prob(x, means, v, w)
    elif kind == "diag":
        var = rs.uniform(0.03, 0.3, size=(K, D))
        mix = _device.DeviceMixture.from_params(means, var, w, _lib.COV_DIAG)
        ref = O.gmm_log_prob(x, means, var, w)
    else:
        covs = np.empty((K, D, D))
        for k in range(K):
            a = rs.normal(size=(D, D))
            covs[k] = 0.1 * a @ a.T / D + 0.05 * np.eye(D)
        mix = _device.DeviceMixture.from_params(means, covs, w, _lib.COV_FULL)
        ref = O.mixture_log_prob(x, means, covs, w)
    got = _device.mixture_logpdf(torch.from_numpy(x).cuda().float(), mix).cpu().numpy()
    assert rel_err(got, ref) < LOGP_TOL


@pytest.mark.parametrize("n,K,D", [(300, 5000, 8), (1, 4100, 2), (2049, 3000, 3), (513, 2048, 32)])
def test_split_component_path_vs_oracle(n, K, D):
    """Few points x many centres takes the split-K + merge kernels (workspace path)."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(n + K)
    centres = rs.normal(size=(K, D)) * 0.4
    x = rs.normal(size=(n, D)) * 0.5
    bw = 0.02
    mix = _device.DeviceMixture.kde(torch.from_numpy(centres).cuda().float(), None, K, bw)
    got = _device.mixture_logpdf(torch.from_numpy(x).cuda().float(), mix).cpu().numpy()
    ref = O.iso_mixture_log_prob(x.astype(np.float32).astype(np.float64), centres.astype(np.float32).astype(np.float64), bw)
    assert rel_err(got, ref) < LOGP_TOL


def test_special_values():
    from ais_benchmarks_b200 import _device
    g = load_golden("gmm")
    dist = gmm(g["means3"], g["sigmas3"], g["weights3"])
    x = g["x3"][:8].copy()
    x[1, 2] = np.nan
    x[2, 0] = np.inf
    x[3, :] = 1e30
    lp = dist.log_prob(x)
    # NaN in -> NaN out; +-inf coordinates give NaN or -inf (the reference gives NaN for D > 1: inf * 0 in its matmul);
    # huge finite coordinates underflow to -inf
    assert np.isnan(lp[1]) and (np.isnan(lp[2]) or np.isneginf(lp[2])) and np.isneginf(lp[3])
    assert rel_err(lp[[0, 4, 5, 6, 7]], g["logp3"][[0, 4, 5, 6, 7]]) < LOGP_TOL
    assert dist.log_prob(np.zeros((0, 8))).shape == (0,)
    assert _device.mixture_logpdf(torch.zeros((0, 8), device="cuda"), dist.device_mixture()).shape == (0,)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", range(3))
def test_dm_weights_golden(case):
    """Reference's own samples -> fused weight kernel; normalised weights and NESS vs the per-sample loop."""
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
    g = load_golden("dm_weights")
    x, pm, sigma = g["x%d" % case], g["pmeans%d" % case], float(g["sigma%d" % case])
    target = gmm(g["tmeans%d" % case], g["tsigmas%d" % case], g["tweights%d" % case])
    D = x.shape[1]
    s = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D,
                                  "K": 1, "N": len(pm), "J": 1, "sigma": sigma})
    s._set_means_host(pm)
    xd = torch.from_numpy(x).cuda().float()
    logw, partials = s.weigh(xd, target)
    ref_w = g["w%d" % case]
    fin = np.isfinite(ref_w) & (ref_w > 0)
    assert rel_err(logw.cpu().numpy()[fin], np.log(ref_w[fin])) < 5e-5  # difference of two 1e-5 log densities
    s._append(xd, logw)
    from ais_benchmarks_b200 import _device
    nw = _device.normalize_weights(logw, s._weight_partials()).cpu().numpy()
    ref_nw = g["norm_w%d" % case]
    big = ref_nw > 1e-12
    assert np.max(np.abs(nw[big] - ref_nw[big]) / ref_nw[big]) < W_TOL
    assert abs(nw.sum() - 1.0) < 1e-5
    assert s.get_NESS() == pytest.approx(float(g["ness%d" % case]), rel=W_TOL)
    # the sampler's own log_prob is the pre-adaptation proposal mixture
    assert rel_err(s.log_prob(x), g["logq%d" % case]) < LOGP_TOL
    # target given only as a foreign object with prob/log_prob (numpy on the host) takes the logp_in path
    class Foreign:
        def log_prob(self, v):
            return O.gmm_log_prob(v, g["tmeans%d" % case], g["tsigmas%d" % case], g["tweights%d" % case])
    logw2, _ = s.weigh(xd, Foreign())
    assert rel_err(logw2.cpu().numpy()[fin], np.log(ref_w[fin])) < 5e-5


def test_dm_weights_unfused_paths():
    """Full-covariance target and split-K proposals go through the unfused branch of aisb_dm_weights_f32."""
    from ais_benchmarks_b200 import _device, _lib
    from ais_benchmarks_b200.distributions import CMultivariateNormal
    rs = np.random.RandomState(5)
    D = 4
    a = rs.normal(size=(D, D))
    cov = a @ a.T / D + 0.2 * np.eye(D)
    tgt = CMultivariateNormal({"mean": np.zeros(D), "sigma": cov})
    for N, n in [(16, 1000), (4096, 600)]:
        pm = rs.normal(size=(N, D))
        x = rs.normal(size=(n, D))
        q = _device.DeviceMixture.from_params(pm, np.array([0.3]), None, _lib.COV_ISO_SHARED)
        logw, logp, logq, part = _device.dm_weights(torch.from_numpy(x).cuda().float(), tgt._device_mixture(), None, q,
                                                    want_logp=True, want_logq=True)
        ref_p = O.mvn_log_prob(x, np.zeros(D), cov)
        ref_q = O.iso_mixture_log_prob(x, pm, 0.3)
        assert rel_err(logp.cpu().numpy(), ref_p) < LOGP_TOL
        assert rel_err(logq.cpu().numpy(), ref_q) < LOGP_TOL
        p = part.cpu().numpy()
        ref_lw = ref_p - ref_q
        assert p[3] == n
        assert p[0] + np.log(p[1]) == pytest.approx(np.log(np.exp(ref_lw - ref_lw.max()).sum()) + ref_lw.max(), abs=1e-4)


def test_device_resident_records_and_streamed_weighting():
    """Device-side record merge + normalisation (no host round trip) and the chunked host->device streamed weighting
    give the values of the host-merged path; checked against the float64 oracle too."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(11)
    for D, N, per in [(2, 64, 400), (32, 128, 700), (8, 33, 129)]:
        mu = rs.uniform(-0.7, 0.7, size=(4, D))
        sig = rs.uniform(0.04, 0.05, size=(4, D))
        w = np.array([0.1, 0.2, 0.3, 0.4])
        target = gmm(mu, sig, w)
        pm = rs.uniform(-1, 1, size=(N, D))
        hint_h = np.repeat(np.arange(N), per).astype(np.int32)
        x = (pm[hint_h] + rs.standard_normal((N * per, D)) * np.sqrt(0.05)).astype(np.float32)
        n = len(x)
        pm_d = torch.from_numpy(pm).cuda().float().contiguous()
        q = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
        tmix = target.device_mixture()
        xd = torch.from_numpy(x).cuda()
        hint = torch.from_numpy(hint_h).cuda()
        logw, _, _, part = _device.dm_weights(xd, tmix, None, q, hint=hint)
        ph = part.cpu().numpy()
        w_host = _device.normalize_weights(logw, ph).cpu().numpy()
        # (1) record in HBM -> same weights, statistics = host helpers
        w_dev, stats = _device.normalize_weights_dev(logw, part)
        st = stats.cpu().numpy()
        assert np.array_equal(w_dev.cpu().numpy(), w_host)
        assert st[0] == pytest.approx(_device.weight_logsum(ph), rel=1e-12)
        assert st[1] == pytest.approx(_device.weight_ess(ph), rel=1e-12)
        assert st[2] == n
        # an odd (4-byte aligned) view takes the scalar branch of the same kernel
        w_odd, _ = _device.normalize_weights_dev(logw[1:], part, out=torch.empty(n + 2, device="cuda")[3:])
        assert np.array_equal(w_odd.cpu().numpy(), w_host[1:])
        # (2) records of row shards merged on the device = host merge rule = the one-shot record
        cuts = [0, n // 3 // 4 * 4, n // 2 // 4 * 4, n]
        recs = torch.stack([_device.logw_partials(logw[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
        merged = _device.merge_weight_records(recs).cpu().numpy()
        rh = recs.cpu().numpy()
        host = _device.weight_merge(_device.weight_merge(rh[0], rh[1]), rh[2])
        assert merged[0] == host[0] and merged[3] == host[3] == n
        assert merged[1] == pytest.approx(host[1], rel=1e-12) and merged[2] == pytest.approx(host[2], rel=1e-12)
        assert merged[0] + np.log(merged[1]) == pytest.approx(ph[0] + np.log(ph[1]), abs=1e-6)
        # (3) streamed from pinned host memory in 3 chunks: bit-identical log weights, same normaliser
        xp = torch.from_numpy(x).pin_memory()
        chunk = ((n + 2) // 3 + 1023) // 1024 * 1024
        xs, logw_s, rec_s = _device.dm_weights_streamed(xp, tmix, q, hint=hint, chunk_rows=chunk)
        assert torch.equal(xs, xd)
        assert torch.equal(logw_s, logw)
        assert rec_s.shape == (3, 4)
        rs_ = _device.merge_weight_records(rec_s).cpu().numpy()
        assert rs_[3] == n and rs_[0] + np.log(rs_[1]) == pytest.approx(ph[0] + np.log(ph[1]), abs=1e-6)
        w_s, st_s = _device.normalize_weights_dev(logw_s, rec_s)      # the kernel folds the chunk records itself
        assert st_s.cpu().numpy()[2] == n and st_s.cpu().numpy()[1] == pytest.approx(_device.weight_ess(ph), rel=1e-9)
        assert np.max(np.abs(w_s.cpu().numpy() - w_host)) <= 2e-6 * w_host.max()
        # (4) float64 oracle on the same float32 samples
        x64 = x.astype(np.float64)
        ref_lw = O.gmm_log_prob(x64, mu, sig, w) - O.iso_mixture_log_prob(x64, pm_d.double().cpu().numpy(), 0.05)
        ref_w = np.exp(ref_lw - ref_lw.max())
        ref_w /= ref_w.sum()
        big = ref_w > 1e-9
        assert np.max(np.abs(w_s.cpu().numpy()[big] - ref_w[big]) / ref_w[big]) < W_TOL
    # KLD records on the device
    lp = torch.from_numpy(rs.normal(size=5000).astype(np.float32)).cuda()
    lq = torch.from_numpy(rs.normal(size=5000).astype(np.float32)).cuda()
    recs = torch.stack([_device.kld_partials(lp[:2000], lq[:2000]), _device.kld_partials(lp[2000:], lq[2000:])])
    merged = _device.merge_kld_records(recs).cpu().numpy()
    whole = _device.kld_partials(lp, lq).cpu().numpy()
    assert _device.kld_finalize(merged) == pytest.approx(_device.kld_finalize(whole), rel=1e-9)
    assert _device.kld_finalize(merged) == pytest.approx(
        O.kld_from_log_probs(lp.double().cpu().numpy(), lq.double().cpu().numpy()), rel=KLD_TOL)


@pytest.mark.parametrize("case", range(3))
def test_kde_kld_golden(case):
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.metrics import CKLDivergence
    from ais_benchmarks_b200.sampling_methods.base import CDeviceKDE
    g = load_golden("kde_kld")
    target = gmm(g["tmeans%d" % case], g["tsigmas%d" % case], g["tweights%d" % case])
    centres, bw, x = g["centres%d" % case], float(g["bw%d" % case]), g["x%d" % case]
    D = x.shape[1]
    kde = CDeviceKDE(torch.from_numpy(centres).cuda().float(), None, len(centres), bw, D, [x.min(0), x.max(0)])
    assert rel_err(kde.log_prob(x), g["lq%d" % case]) < LOGP_TOL
    assert rel_err(target.log_prob(x), g["lp%d" % case]) < LOGP_TOL
    metric = CKLDivergence()
    val = metric.compute_from_samples(target, kde, x)
    ref = float(g["kld%d" % case])
    assert abs(val - ref) <= W_TOL * max(1.0, abs(ref))
    assert metric.value == val
    # per-inlier terms + in-place normalisation of the inputs (quirk Q7)
    lp, lq = g["lp%d" % case].copy(), g["lq%d" % case].copy()
    terms = CKLDivergence.compute_from_log_probs(lp, lq)
    assert abs(terms.sum() - ref) <= W_TOL * max(1.0, abs(ref))
    inl = np.isfinite(g["lp%d" % case]) & np.isfinite(g["lq%d" % case])
    assert abs(np.logaddexp.reduce(lp[inl])) < 1e-4
    # foreign q (numpy-only object) still works
    class ForeignQ:
        def log_prob(self, v):
            return O.iso_mixture_log_prob(v, centres, bw)
    val2 = CKLDivergence().compute_from_samples(target, ForeignQ(), x)
    assert abs(val2 - ref) <= W_TOL * max(1.0, abs(ref))


@pytest.mark.parametrize("case", range(3))
def test_jsd_evmse_golden(case):
    """CJSDivergence (divergences.py:156-195) and CExpectedValueMSE (statistics.py:29-48) on the KDE fixtures against
    the values the reference returned for the same p, q and evaluation points."""
    from ais_benchmarks_b200.metrics import CExpectedValueMSE, CJSDivergence
    from ais_benchmarks_b200.sampling_methods.base import CDeviceKDE
    g, k = load_golden("metrics"), load_golden("kde_kld")
    target = gmm(k["tmeans%d" % case], k["tsigmas%d" % case], k["tweights%d" % case])
    centres, bw, x = k["centres%d" % case], float(k["bw%d" % case]), k["x%d" % case]
    kde = CDeviceKDE(torch.from_numpy(centres).cuda().float(), None, len(centres), bw, x.shape[1], [x.min(0), x.max(0)])
    jsd = CJSDivergence()
    val = jsd.compute_from_samples(target, kde, x)
    ref = float(g["jsd%d" % case])
    assert abs(val - ref) <= W_TOL * max(1.0, abs(ref)), (val, ref)
    assert jsd.value == val and jsd.dropped == 0
    ev = CExpectedValueMSE()
    val = ev.compute_from_samples(target, kde, torch.from_numpy(x).cuda())  # device points in, same statistic
    ref = float(g["evmse%d" % case])
    assert abs(val - ref) <= W_TOL * max(1.0, abs(ref)), (val, ref)
    assert np.allclose(ev.ev_p, g["ev_p%d" % case], atol=1e-4) and np.allclose(ev.ev_q, g["ev_q%d" % case], atol=1e-4)


def test_jsd_evmse_edge_cases_and_compute():
    from ais_benchmarks_b200.distributions import CMultivariateNormal, CMultivariateUniform
    from ais_benchmarks_b200.metrics import CExpectedValueMSE, CJSDivergence
    g = load_golden("metrics")
    with np.errstate(all="ignore"):
        lp = torch.from_numpy(np.log(g["edge_p"])).cuda().float()
        lq = torch.from_numpy(np.log(g["edge_q"])).cuda().float()
        fp = torch.from_numpy(np.log(g["edge_fp"])).cuda().float()
        fq = torch.from_numpy(np.log(g["edge_fq"])).cuda().float()
    # zero / NaN probabilities on either side are not inliers (divergences.py:170-173)
    assert CJSDivergence().compute_from_device_log_probs(lp, lq) == pytest.approx(float(g["edge_jsd"]), rel=1e-5)
    ex = torch.from_numpy(g["edge_x"]).cuda().float()
    assert CExpectedValueMSE().compute_from_device_log_probs(ex, fp, fq) == pytest.approx(float(g["edge_evmse"]), rel=1e-4)
    # empty inlier set -> 0.0; a NaN density -> NaN expectation, as in the reference
    none = torch.full((5,), -np.inf, device="cuda")
    assert CJSDivergence().compute_from_device_log_probs(none, torch.zeros(5, device="cuda")) == 0.0
    assert np.isnan(CExpectedValueMSE().compute_from_device_log_probs(ex, lp[:6].contiguous(), fq))
    # compute(): seeded host evaluation points reproduce the reference's value on the same points
    p = CMultivariateNormal({"mean": np.array([0.0]), "sigma": np.array([[1.0]])})
    q = CMultivariateNormal({"mean": np.array([1.0]), "sigma": np.array([[1.0]])})
    np.random.seed(70)
    assert CJSDivergence().compute(p=p, q=q, nsamples=20000) == pytest.approx(float(g["norm_jsd"]), abs=W_TOL)
    np.random.seed(70)
    assert CExpectedValueMSE().compute(p=p, q=q, nsamples=20000) == pytest.approx(float(g["norm_evmse"]), rel=W_TOL)
    assert abs(CJSDivergence().compute(p=p, q=p, nsamples=5000)) < 1e-6
    # disjoint supports: no inliers -> 0.0 (the reference warns and returns the empty sum)
    pu = CMultivariateUniform({"center": np.array([0.0]), "radius": np.array([.1])})
    qu = CMultivariateUniform({"center": np.array([1.0]), "radius": np.diag([.1])})
    assert CJSDivergence().compute(p=pu, q=qu, nsamples=1000) == 0.0
    with pytest.raises(AssertionError):
        CJSDivergence().compute(p=p, q=object(), nsamples=10)
    # D that does not divide the block (weighted_mean_kernel keeps (256 / D) * D threads active), large n
    rs = np.random.RandomState(9)
    for D in (1, 3, 7, 64):
        x = rs.normal(size=(40001, D)).astype(np.float32)
        lw = rs.normal(size=40001).astype(np.float32) * 3
        from ais_benchmarks_b200 import _device
        got = _device.weighted_mean(torch.from_numpy(x).cuda(), torch.from_numpy(lw).cuda(), 0.25).cpu().numpy()
        ref = (x.astype(np.float64) * np.exp(lw.astype(np.float64) - 0.25)[:, None]).sum(0)
        assert np.allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_kld_edge_cases_and_known_answers():
    from ais_benchmarks_b200.distributions import CMultivariateNormal, CMultivariateUniform
    from ais_benchmarks_b200.metrics import CKLDivergence
    g = load_golden("kde_kld")
    terms = CKLDivergence.compute_from_log_probs(g["edge_lp"].copy(), g["edge_lq"].copy())
    assert terms.sum() == pytest.approx(float(g["edge_kld"]), rel=1e-5)
    assert np.sum(CKLDivergence.compute_from_log_probs(np.full(4, -np.inf), np.zeros(4))) == 0.0
    # reference tests/metrics/test_metric_kldivergence.py:77-135 (closed-form Gaussian KL within 1 %)
    for j in range(int(g["n_kat"])):
        m0, m1, s0, s1 = g["kat_m0_%d" % j], g["kat_m1_%d" % j], g["kat_s0_%d" % j], g["kat_s1_%d" % j]
        p = CMultivariateNormal({"mean": m0, "sigma": s0})
        q = CMultivariateNormal({"mean": m1, "sigma": s1})
        np.random.seed(int(g["kat_seed_%d" % j]))
        val = CKLDivergence().compute(p=p, q=q, nsamples=100000)
        ref = float(g["kat_kld_%d" % j])
        assert abs(val - ref) <= W_TOL * max(1.0, abs(ref))
        assert abs(val - O.mvn_kld_closed_form(m0, s0, m1, s1)) <= 0.01 * O.mvn_kld_closed_form(m0, s0, m1, s1)
    # test_equal: identical distributions -> 0; test_disjoint: disjoint uniforms -> empty inlier set -> 0.0
    p = CMultivariateNormal({"mean": np.array([0.0]), "sigma": np.diag([1.0])})
    assert abs(CKLDivergence().compute(p=p, q=p, nsamples=10000)) < 1e-6
    pu = CMultivariateUniform({"center": np.array([0.0]), "radius": np.array([.1])})
    qu = CMultivariateUniform({"center": np.array([1.0]), "radius": np.diag([.1])})
    assert CKLDivergence().compute(p=pu, q=qu, nsamples=1000) == 0.0
    # device-generated evaluation points: same statistic within Monte-Carlo noise
    m = CKLDivergence()
    m.device_points = True
    q2 = CMultivariateNormal({"mean": np.array([2.0]), "sigma": np.diag([1.0])})
    assert m.compute(p=p, q=q2, nsamples=200000) == pytest.approx(2.0, rel=0.02)


def test_mpmc_iteration_golden():
    from ais_benchmarks_b200.sampling_methods import CMixturePMC
    g = load_golden("mpmc")
    for i in range(int(g["n"])):
        target = gmm(g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i])
        for it in range(2):
            t = "%d_%d" % (i, it)
            x = g["x" + t]
            K, D = x.shape
            N = len(g["means" + t])
            np.random.seed(0)
            s = CMixturePMC({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": K, "N": N,
                             "J": 1000, "sigma": float(g["sigma%d" % i])})
            s._means, s._covs = g["means" + t].copy(), g["covs" + t].copy()
            s.wproposals = g["alpha" + t].copy()
            s._mix_weights = g["mixw" + t].copy()
            s._mixture = None
            xd = torch.from_numpy(x).cuda().float()
            logw, logq, part = s.weigh(xd, target)
            ph = part.cpu().numpy()
            from ais_benchmarks_b200 import _device
            w = torch.exp(logw.double() - _device.weight_logsum(ph)).cpu().numpy()
            ref_w = g["w" + t]
            big = ref_w > 1e-9
            assert np.max(np.abs(w[big] - ref_w[big]) / ref_w[big]) < W_TOL
            pre = (s._means.copy(), s._covs.copy(), s.wproposals.copy(), s._mix_weights.copy())
            s.adapt(xd, logw, logq, ph)
            assert np.allclose(s.wproposals, g["new_alpha" + t], rtol=W_TOL, atol=1e-9)
            live = g["new_alpha" + t] > 1e-9
            assert np.allclose(s._means[live], g["new_means" + t][live], rtol=W_TOL, atol=1e-5)
            assert np.allclose(s._covs[live], g["new_covs" + t][live], rtol=1e-3, atol=1e-5)
            # the same update by the device kernel (aisb_mpmc_update_f64): must agree with the host float64 update
            host = (s._means.copy(), s._covs.copy(), s.wproposals.copy(), s._mix_weights.copy())
            s._means, s._covs = pre[0], pre[1]
            s.wproposals, s._mix_weights = pre[2], pre[3]
            logw2, logq2, part2 = s.weigh(xd, target)
            s.adapt_device(xd, logw2, logq2, part2, K)
            ok = np.isfinite(host[0]).all(axis=1)
            assert np.allclose(s._means[ok], host[0][ok], rtol=1e-9, atol=1e-12)
            assert np.allclose(s._covs[ok], host[1][ok], rtol=1e-9, atol=1e-12)
            assert np.allclose(s.wproposals, host[2], rtol=1e-9, atol=1e-15, equal_nan=True)
            assert np.allclose(s._mix_weights, host[3], rtol=1e-9, atol=1e-15)
            # ... and the mixture it packed on the device evaluates like the one the host packs from the same parameters
            lq_dev = _device.mixture_logpdf(xd, s.proposal_mixture()).cpu().numpy()
            s._means = s._means                                 # hand the parameters back to the host packer
            lq_host = _device.mixture_logpdf(xd, s.proposal_mixture()).cpu().numpy()
            assert rel_err(lq_dev, lq_host.astype(np.float64)) < 2e-6, rel_err.last


@pytest.mark.parametrize("D", [1, 3, 8, 16, 32])
@pytest.mark.parametrize("kind", ["iso", "diag", "full"])
def test_mpmc_moment_kernels_vs_oracle(kind, D):
    """Rao-Blackwellised M-PMC moments (m_pmc.py:57-61,102-105) for every kernel variant -- the component-stationary
    kernel (small rows) and the point-stationary one (large rows) -- against a float64 evaluation of
    sum_i w~_i rho_di [1, x_i]; ragged sizes (n not a multiple of the tiles, K not a multiple of 128)."""
    from ais_benchmarks_b200 import _device, _lib
    from scipy.special import logsumexp
    rs = np.random.RandomState(100 * D + len(kind))
    n, K = 2048 * 2 + 517, 150
    mu = rs.uniform(-1, 1, size=(K, D))
    w = rs.rand(K)
    w[5] = 0.0
    w /= w.sum()
    x = rs.uniform(-1.2, 1.2, size=(n, D)).astype(np.float32)
    x64 = x.astype(np.float64)
    if kind == "iso":
        covs = np.stack([np.eye(D) * 0.3] * K)
        mix = _device.DeviceMixture.from_params(mu, np.array([0.3]), w, _lib.COV_ISO_SHARED)
    elif kind == "diag":
        var = rs.uniform(0.1, 0.4, size=(K, D))
        covs = np.stack([np.diag(v) for v in var])
        mix = _device.DeviceMixture.from_params(mu, var, w, _lib.COV_DIAG)
    else:
        a = rs.normal(size=(K, D, D)) * 0.2
        covs = a @ np.transpose(a, (0, 2, 1)) + 0.15 * np.eye(D)
        mix = _device.DeviceMixture.from_params(mu, covs, w, _lib.COV_FULL)
    with np.errstate(divide="ignore"):
        comp = np.stack([O.mvn_log_prob(x64, mu[k], covs[k]) + np.log(w[k]) for k in range(K)])   # [K, n]
    logq = logsumexp(comp, axis=0)
    logw = rs.normal(size=n) * 2.0
    logw[7] = -np.inf                     # a zero-weight sample contributes nothing
    log_norm = logsumexp(logw)
    a_ki = np.exp(comp - logq[None, :] + (logw - log_norm)[None, :])
    ref = np.concatenate([a_ki.sum(1, keepdims=True), a_ki @ x64], axis=1)
    xd = torch.from_numpy(x).cuda()
    got = _device.mpmc_moments(xd, mix, torch.from_numpy(logq).cuda().float(), torch.from_numpy(logw).cuda().float(),
                               float(log_norm)).cpu().numpy()
    assert got.shape == (K, D + 1)
    assert np.all(got[5] == 0.0)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-12
    assert np.max(np.abs(got - ref) / scale) < W_TOL
    assert abs(got[:, 0].sum() - 1.0) < 1e-4      # sum_d alpha_d = 1
    if D <= 8:
        # weighted second moments (C ABI aisb_mpmc_moments2_f32): lower triangle of sum_i a_di x_i x_i^T
        got2 = _device.mpmc_moments(xd, mix, torch.from_numpy(logq).cuda().float(),
                                    torch.from_numpy(logw).cuda().float(), float(log_norm), second=True).cpu().numpy()
        assert got2.shape == (K, D + 1 + D * (D + 1) // 2)
        assert np.max(np.abs(got2[:, :D + 1] - ref) / scale) < W_TOL
        r, c = np.tril_indices(D)
        ref2 = np.einsum("ki,ir,ic->krc", a_ki, x64, x64)[:, r, c]
        scale2 = np.abs(ref2).max(axis=1, keepdims=True) + 1e-12
        assert np.max(np.abs(got2[:, D + 1:] - ref2) / scale2) < W_TOL
    else:
        with pytest.raises(_lib.AisbError):
            _device.mpmc_moments(xd, mix, torch.from_numpy(logq).cuda().float(), torch.from_numpy(logw).cuda().float(),
                                 float(log_norm), second=True)


def test_mpmc_paper_covariance_adapts():
    """CMixturePMC(paper_covariance=True): the weighted covariance update of the M-PMC paper (eq. 14.3) instead of the
    reference's unweighted scatter (quirk Q3, which always ends in the 10 I failsafe). One adapt step is checked against
    a float64 evaluation of the update; over a few iterations the proposal mixture approaches the target."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.metrics import CKLDivergence
    from ais_benchmarks_b200.sampling_methods import CMixturePMC
    from scipy.special import logsumexp
    rs = np.random.RandomState(12)
    D, N, K = 2, 6, 20000
    mu = np.array([[-0.5, -0.4], [0.5, 0.3], [0.0, 0.6]])
    sig = rs.uniform(0.01, 0.03, size=(3, D))
    w = np.array([0.3, 0.5, 0.2])
    sup = [np.full(D, -1.0), np.full(D, 1.0)]
    target = gmm(mu, sig, w, support=sup)
    params = {"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 10, "sigma": 0.05,
              "paper_covariance": True}
    np.random.seed(5)
    _device.seed(5)
    s = CMixturePMC(params)
    # one adapt step on a fixed batch vs float64
    x = s.sample_batch()
    logw, logq, part = s.weigh(x, target)
    m0, c0, a0 = s._means.copy(), s._covs.copy(), s._mix_weights.copy()
    s.adapt(x, logw, logq, part.cpu().numpy())
    x64 = x.double().cpu().numpy()
    comp = np.stack([O.mvn_log_prob(x64, m0[k], c0[k]) + np.log(a0[k]) for k in range(N)])
    lq = logsumexp(comp, axis=0)
    lw = O.gmm_log_prob(x64, mu, sig, w) - lq
    wt = np.exp(lw - logsumexp(lw))
    a = wt[None, :] * np.exp(comp - lq[None, :])
    alpha = a.sum(1)
    mean = (a @ x64) / alpha[:, None]
    cov = np.einsum("ki,kir,kic->krc", a, x64[None] - mean[:, None], x64[None] - mean[:, None]) / alpha[:, None, None]
    live = alpha > 1e-6
    assert np.allclose(s.wproposals, alpha / alpha.sum(), rtol=1e-3, atol=1e-7)
    assert np.allclose(s._means[live], mean[live], rtol=1e-3, atol=1e-4)
    assert np.allclose(s._covs[live], cov[live], rtol=2e-2, atol=2e-5)
    # a few more iterations: the proposal mixture closes in on the target, covariances at the target's scale
    xe = rs.uniform(-1, 1, size=(20000, D))
    kld = CKLDivergence()
    np.random.seed(6)
    s2 = CMixturePMC(params)
    k0 = kld.compute_from_samples(target, s2, xe)
    s2.importance_sample(target, 6 * K)
    k1 = kld.compute_from_samples(target, s2, xe)
    assert k1 < 0.25 * k0 and k1 < 0.5, (k0, k1)
    heavy = s2._mix_weights > 0.05
    assert np.all(np.abs(s2._covs[heavy]) < 0.2)
    with pytest.raises(NotImplementedError):
        CMixturePMC(dict(params, dims=9, space_min=np.full(9, -1.0), space_max=np.full(9, 1.0)))


def test_heterogeneous_mixture_golden():
    """CMixtureModel over Gaussian + uniform components (the reference's loop takes any models[], CMixtureModel.py:53-56;
    round 1 raised NotImplementedError): packed Gaussian sub-mixture + box sub-mixture joined by a device log-sum-exp,
    against the reference's values (tests/golden/make_golden.py::make_hetero). A zero-weight component contributes
    nothing; points outside every box and far from every Gaussian keep a finite density from the Gaussians."""
    from ais_benchmarks_b200.distributions import CMixtureModel, CMultivariateNormal, CMultivariateUniform
    g = load_golden("hetero")
    comps = [CMultivariateNormal({"mean": np.array([0.3, -0.2]), "sigma": np.diag([0.05, 0.02])}),
             CMultivariateUniform({"center": np.array([-0.4, 0.1]), "radius": np.array([0.3, 0.2])}),
             CMultivariateNormal({"mean": np.array([-0.5, 0.5]), "sigma": np.array([[0.04, 0.01], [0.01, 0.03]])}),
             CMultivariateUniform({"center": np.array([0.5, 0.5]), "radius": np.array([0.25])}),
             CMultivariateNormal({"mean": np.array([0.9, 0.9]), "sigma": np.diag([0.01, 0.01])})]
    mix = CMixtureModel({"models": comps, "weights": g["w"].copy(), "dims": 2, "support": [[-1, -1], [1, 1]]})
    assert rel_err(mix.log_prob(g["x"]), g["logp"]) < LOGP_TOL, rel_err.last
    assert np.allclose(mix.prob(g["x"]), g["prob"], rtol=1e-5, atol=1e-300)
    s = mix.sample(2000)
    assert s.shape == (2000, 2) and np.isfinite(mix.log_prob(s)).all()
    in_box = np.all(np.abs(s - np.array([-0.4, 0.1])) <= np.array([0.3, 0.2]), axis=1).mean()
    assert 0.15 < in_box < 0.40           # ~0.2 from the first box + the Gaussian mass that overlaps it


def test_lais_chains_golden():
    from ais_benchmarks_b200.sampling_methods import CLayeredAIS
    g = load_golden("lais")
    for i in range(int(g["n"])):
        target = gmm(g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i])
        before, delta, u = g["before%d" % i], g["delta%d" % i], g["u%d" % i]
        N, D = before.shape
        np.random.seed(1)
        s = CLayeredAIS({"space_min": g["lo%d" % i], "space_max": g["hi%d" % i], "dims": D, "K": 2, "N": N, "J": 1,
                         "L": delta.shape[0], "sigma": 0.05, "mh_sigma": 0.5})
        s._set_means_host(before)
        s.mcmc_step(target, torch.from_numpy(delta).cuda().float(), torch.from_numpy(u).cuda().float())
        after = s.proposal_means().cpu().numpy()
        assert np.allclose(after, g["after%d" % i], rtol=0, atol=2e-6)
        assert s.num_target_evals == int(g["pi_evals%d" % i])


def test_uniform_and_box_mixture_golden():
    from ais_benchmarks_b200.distributions import CMixtureModel, CMultivariateUniform
    g = load_golden("uniform")
    for i in range(int(g["n"])):
        d = CMultivariateUniform({"center": g["c%d" % i].copy(), "radius": g["r%d" % i].copy()})
        x = g["x%d" % i]
        assert rel_err(d.log_prob(x), g["logp%d" % i]) < LOGP_TOL
        assert np.allclose(d.prob(x), g["prob%d" % i], rtol=1e-5)
    c, r = g["mix_centers"], g["mix_radii"]
    models = [CMultivariateUniform({"center": c[j].copy(), "radius": r[j].copy()}) for j in range(len(c))]
    mix = CMixtureModel({"models": models, "weights": g["mix_w"].copy(), "dims": 2, "support": [[-2, -2], [2, 2]]})
    assert rel_err(mix.log_prob(g["mix_x"]), g["mix_logp"]) < LOGP_TOL
    s = mix.sample(500)
    assert s.shape == (500, 2) and np.isfinite(mix.log_prob(s)).all()


# ---------------------------------------------------------------------------------------------------
def test_samplers_end_to_end():
    """importance_sample / adapt loops of DM-AIS, LAIS and M-PMC on the default 2-D GMM target
    (benchmark/def_benchmark.yaml:15-20): interface, counters and weight normalisation contracts."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.metrics import CKLDivergence
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS, CMixturePMC
    np.random.seed(0)
    _device.seed(0)
    target = gmm(np.array([[0.0, 0.0], [-0.3, 0.5]]), np.array([[0.01, 0.1], [0.01, 0.03]]), np.array([0.5, 0.5]),
                 [[-1, -1], [1, 1]])
    base = {"space_min": target.support()[0], "space_max": target.support()[1], "dims": 2}
    dm = CDeterministicMixtureAIS(dict(base, K=10, N=5, J=1000, sigma=0.5))
    x, w = dm.importance_sample(target, 120, timeout=60)
    assert x.shape == (150, 2) and w.shape == (150,) and abs(w.sum() - 1) < 1e-5
    assert dm.num_proposal_samples == 150 and dm.num_target_evals == 150
    assert 0 < dm.get_NESS() <= 1 and dm.get_acceptance_rate() == 1.0
    x2, w2 = dm.importance_sample(target, 160, timeout=60)  # cumulative, like CBenchmark.py:489-493
    assert len(x2) == 200 and np.array_equal(x2[:150], x)
    assert len(dm.proposals) == 5 and dm.model.log_prob(x2[:3]).shape == (3,)
    assert rel_err(dm.log_prob(x2[:50]), dm.model.log_prob(x2[:50])) < LOGP_TOL
    kld = CKLDivergence().compute(p=target, q=dm, nsamples=2000)
    assert np.isfinite(kld) and kld > 0
    dm.reset()
    assert dm.num_proposal_samples == 0 and len(dm.samples) == 0

    lais = CLayeredAIS(dict(base, K=3, N=5, J=1000, L=10, sigma=0.5, mh_sigma=0.5))
    x, w = lais.importance_sample(target, 40, timeout=60)
    assert x.shape == (45, 2) and abs(w.sum() - 1) < 1e-5
    assert lais.num_target_evals == 45 + 3 * 2 * 10 * 5 and lais.num_proposal_samples == 45 + 3 * 10 * 5
    m = lais.proposal_means().cpu().numpy()
    assert np.all(m >= -1) and np.all(m <= 1)

    pmc = CMixturePMC(dict(base, K=50, N=10, J=1000, sigma=0.01))
    x, w = pmc.importance_sample(target, 200, timeout=60)
    assert x.shape == (200, 2) and w.sum() == pytest.approx(4.0, rel=1e-4)  # quirk Q4: one unit of weight per batch
    assert pmc.num_target_evals == 200 and pmc.num_proposal_evals == 200 * 11
    assert len(pmc.proposals) == 10 and abs(pmc.wproposals.sum() - 1) < 1e-9
    assert np.isfinite(CKLDivergence().compute(p=target, q=pmc, nsamples=2000))
    # tensors out on request
    dmt = CDeterministicMixtureAIS(dict(base, K=10, N=5, J=1000, sigma=0.5, device_outputs=True))
    xt, wt = dmt.importance_sample(target, 50)
    assert xt.is_cuda and wt.is_cuda and abs(float(wt.sum()) - 1) < 1e-5


def test_kde_base_class_builder():
    """CMixtureSamplingMethod._update_model (sampling_methods/base.py:162-179) for any sampler that accumulates samples."""
    from ais_benchmarks_b200.sampling_methods import CMixtureSamplingMethod

    class Holder(CMixtureSamplingMethod):
        def __init__(self, params):
            super().__init__(params)
            self.bw = params["kde_bw"]

        def importance_sample(self, target_d, n_samples, timeout):
            raise NotImplementedError

    rs = np.random.RandomState(3)
    pts = rs.normal(size=(50, 2)) * 0.3
    h = Holder({"space_min": np.array([-1.0, -1]), "space_max": np.array([1.0, 1]), "dims": 2, "n_samples_kde": 50,
                "kde_bw": 0.02})
    h.samples = pts
    np.random.seed(0)
    x = rs.uniform(-1, 1, size=(300, 2))
    lq = h.log_prob(x)
    idx = h.model._idx.cpu().numpy()
    assert idx.shape == (50,) and idx.min() >= 0 and idx.max() < 50
    ref = O.iso_mixture_log_prob(x, pts[idx].astype(np.float32).astype(np.float64), 0.02)
    assert rel_err(lq, ref) < LOGP_TOL
    assert h.sample(7).shape == (7, 2) and h.num_proposal_samples == 7


def test_device_sampling_statistics():
    from ais_benchmarks_b200 import _device
    _device.seed(123)
    means = torch.tensor([[0.0, 10.0, -5.0], [1.0, 2.0, 3.0]], device="cuda")
    x = _device.sample_iso(means, 200000, 0.25)
    assert x.shape == (400000, 3)
    a, b = x[:200000], x[200000:]
    assert torch.allclose(a.mean(0), means[0], atol=0.01) and torch.allclose(b.mean(0), means[1], atol=0.01)
    assert torch.allclose(a.var(0), torch.full((3,), 0.25, device="cuda"), rtol=0.02)
    z = (a - means[0]) / 0.5
    assert abs(float((z ** 4).mean()) - 3.0) < 0.06 and abs(float((z[:, 0] * z[:, 1]).mean())) < 0.01
    u = _device.uniform_box(torch.tensor([-1.0, 2.0], device="cuda"), torch.tensor([1.0, 6.0], device="cuda"), 100000)
    assert float(u[:, 0].min()) > -1 and float(u[:, 0].max()) < 1 and float(u[:, 1].min()) > 2 and float(u[:, 1].max()) < 6
    assert abs(float(u[:, 1].mean()) - 4.0) < 0.02 and abs(float(u[:, 0].var()) - 4.0 / 12) < 0.01
    y = _device.sample_iso(means, 10, 0.25)
    assert not torch.equal(y, x[:20].reshape(-1, 3)[:20])  # the stream advances between calls


@pytest.mark.parametrize("D", [1, 3, 8, 33])
def test_mixture_sampling_kernel_statistics(D):
    """x = mu_k + L_k z draws (C ABI aisb_sample_mixture_f32): per-component mean / covariance, independence of the
    draws, and M-PMC's proposal sampler on top of it (CMixtureModel.sample semantics, CMixtureModel.py:79-89)."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(D)
    K, n = 3, 240000
    mu = rs.uniform(-3, 3, size=(K, D))
    A = rs.normal(size=(K, D, D)) * 0.3
    cov = A @ np.transpose(A, (0, 2, 1)) + 0.2 * np.eye(D)[None]
    L = np.linalg.cholesky(cov)
    _device.seed(7)
    idx = torch.from_numpy(np.repeat(np.arange(K), n // K)).cuda()
    x = _device.sample_mixture(torch.from_numpy(mu).cuda().float(), torch.from_numpy(L).cuda().float().contiguous(), idx)
    assert x.shape == (n, D)
    xs = x.double().cpu().numpy().reshape(K, n // K, D)
    for k in range(K):
        assert np.allclose(xs[k].mean(0), mu[k], atol=0.02 * np.sqrt(np.diag(cov[k]).max()) * 3)
        emp = np.cov(xs[k].T).reshape(D, D)
        assert np.max(np.abs(emp - cov[k])) < 0.03 * np.abs(cov[k]).max() + 0.01
    # whitened draws are standard normal: fourth moment 3, no correlation between consecutive draws
    z = np.linalg.solve(L[0], (xs[0] - mu[0]).T).T
    assert abs((z ** 4).mean() - 3.0) < 0.1
    assert abs((z[1:, 0] * z[:-1, 0]).mean()) < 0.02
    y = _device.sample_mixture(torch.from_numpy(mu).cuda().float(), torch.from_numpy(L).cuda().float().contiguous(), idx)
    assert not torch.equal(x, y)  # the Philox stream advances between calls


def test_mpmc_proposal_sampler():
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CMixturePMC
    np.random.seed(3)
    _device.seed(3)
    D = 4
    s = CMixturePMC({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": 100, "N": 5, "J": 10,
                     "sigma": 0.01, "device_outputs": True})
    s._mix_weights = np.array([0.5, 0.0, 0.25, 0.25, 0.0])
    x = s.sample(200000).double().cpu().numpy()
    # every draw sits within 6 sd of a component with non-zero weight; component frequencies follow the weights
    d2 = ((x[:, None, :] - s._means[None]) ** 2).sum(-1)
    near = d2.argmin(1)
    assert np.all(d2.min(1) < 36 * 0.01 * D)
    freq = np.bincount(near, minlength=5) / len(x)
    assert np.allclose(freq, s._mix_weights, atol=0.01)
    assert s.num_proposal_samples == 200000


def test_sampler_state_growth_buffers():
    """_append keeps the accumulated samples / log weights in capacity-doubling buffers: values, order, and earlier
    views survive many small iterations (the reference's vstack/concatenate growth, dm_ais.py:69,79-80)."""
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
    np.random.seed(0)
    s = CDeterministicMixtureAIS({"space_min": np.full(2, -1.0), "space_max": np.full(2, 1.0), "dims": 2, "K": 1,
                                  "N": 4, "J": 1, "sigma": 0.1})
    rs = np.random.RandomState(0)
    chunks, views = [], []
    for it in range(40):
        k = int(rs.randint(1, 50))
        xs = torch.from_numpy(rs.normal(size=(k, 2)).astype(np.float32)).cuda()
        lw = torch.from_numpy(rs.normal(size=k).astype(np.float32)).cuda()
        s._append(xs, lw)
        chunks.append((xs.cpu().numpy(), lw.cpu().numpy()))
        views.append((s._samples_dev, s._n()))
    allx = np.concatenate([c[0] for c in chunks])
    allw = np.concatenate([c[1] for c in chunks])
    assert np.array_equal(s._samples_dev.cpu().numpy(), allx) and np.array_equal(s._logw_dev.cpu().numpy(), allw)
    assert s._samples_dev.is_contiguous() and s._buf_s.shape[0] >= len(allx)
    for v, n in views[::7]:
        assert np.array_equal(v.cpu().numpy(), allx[:n])
    assert np.allclose(s.samples, allx) and np.allclose(s.weights, np.exp(allw.astype(np.float64)))
    s.weights = np.ones(3)
    s.samples = np.zeros((3, 2))
    s._append(torch.ones((2, 2), device="cuda"), torch.zeros(2, device="cuda"))
    assert s._n() == 5 and np.array_equal(s.samples[3:], np.ones((2, 2)))
    s.reset()
    assert s._n() == 0 and s._buf_s is None


def test_raw_moments():
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(9)
    for n, D in [(1000, 3), (777, 8), (130, 32)]:
        x = rs.normal(size=(n, D)).astype(np.float32)
        s1, s2 = _device.raw_moments(torch.from_numpy(x).cuda())
        xd = x.astype(np.float64)
        assert np.allclose(s1.cpu().numpy(), xd.sum(0), rtol=1e-12, atol=1e-9)
        assert np.allclose(s2.cpu().numpy(), xd.T @ xd, rtol=1e-12, atol=1e-9)


# ---------------------------------------------------------------------------------------------------
def test_large_kde_properties_and_row_subset():
    """BASELINE config-4-shaped run at a size the GPU finishes in well under a second: 2e5 eval points x 2e5 centres,
    D = 8. Checked through (i) an oracle evaluation of a random row subset, (ii) the split/merge identity
    LSE(all centres) = logaddexp(LSE(first half), LSE(second half)) and (iii) invariance to centre permutation."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(4)
    n, D = 200000, 8
    g = load_golden("gmm")
    means, sig, w = g["means3"], g["sigmas3"], g["weights3"]
    comp = rs.choice(len(w), size=n, p=w / w.sum())
    centres = (means[comp] + rs.normal(size=(n, D)) * np.sqrt(sig[comp])).astype(np.float32)
    x = rs.uniform(-1, 1, size=(n, D)).astype(np.float32)
    cd, xd = torch.from_numpy(centres).cuda(), torch.from_numpy(x).cuda()
    bw = 0.02
    full = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd, None, n, bw))
    rows = rs.choice(n, size=48, replace=False)
    ref = O.iso_mixture_log_prob(x[rows].astype(np.float64), centres.astype(np.float64), bw, chunk=8)
    assert rel_err(full.cpu().numpy()[rows], ref) < LOGP_TOL
    h = n // 2
    a = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd[:h].contiguous(), None, h, bw))
    b = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd[h:].contiguous(), None, n - h, bw))
    merged = torch.logaddexp(a.double(), b.double()) + np.log(0.5)
    assert rel_err(full.cpu().numpy(), merged.cpu().numpy()) < 2e-6, rel_err.last
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
    p = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd, perm, n, bw))
    assert rel_err(p.cpu().numpy(), full.cpu().numpy()) < 2e-6, rel_err.last


@pytest.mark.parametrize("D", [1, 2, 3, 8, 16, 32])
def test_morton_key_kernel(D):
    """aisb_morton_keys_f32 against the torch element-wise reference: same codes, a valid permutation, codes
    non-decreasing along it; degenerate columns (zero span) and duplicated rows included."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(D)
    n = 100003
    x = rs.uniform(-3, 7, size=(n, D)).astype(np.float32)
    x[1000:2000] = x[0]                      # duplicates
    if D > 1:
        x[:, D - 1] = 2.5                    # a coordinate without extent
    xd = torch.from_numpy(x).cuda()
    bits = _device.morton_bits(D)
    assert bits * D <= 32
    keys = _device.morton_keys(xd)
    code = (keys.to(torch.int64) & 0xFFFFFFFF) ^ 0x80000000
    lo = xd.min(dim=0).values
    span = (xd.max(dim=0).values - lo).clamp_min(1e-30)
    q = ((xd - lo) / span * float(1 << bits)).clamp_(0, (1 << bits) - 1).to(torch.int64)
    lut = _device._morton_lut(D, bits, xd.device)
    ref = sum(lut[d][q[:, d]] for d in range(D))
    assert float((code == ref).double().mean()) > 0.9999
    perm = _device.morton_order(xd)
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device="cuda"))
    along = code[perm]
    assert bool((along[1:] >= along[:-1]).all())
    perm_t = _device.morton_order_torch(xd)
    assert bool((ref[perm_t][1:] >= ref[perm_t][:-1]).all())


@pytest.mark.parametrize("mode", ["fast", "screen", "ordered"])
@pytest.mark.parametrize("D", [1, 2, 8, 16])
def test_spatially_ordered_kde_path(D, mode, monkeypatch):
    """Large KDEs pack their centres in Z-order and evaluate the points in Z-order with one of three kernel families
    (AISB_KDE_MODE): "fast" = recentred expanded form + exact re-evaluation of the points whose error estimate is too
    large (aisb_mixture_logpdf_fast_f32, the default), "screen" = expanded-form screen + exact refinement
    (aisb_mixture_logpdf_screened_f32), "ordered" = centred form with fold skipping (aisb_mixture_logpdf_ordered_f32).
    Same values as the ordinary path and as the float64 oracle; row order of the output is the caller's; clustered data
    far from the origin included (there the fast kernel's estimate sends most points to the exact re-evaluation)."""
    from ais_benchmarks_b200 import _device
    monkeypatch.setenv("AISB_KDE_MODE", mode)
    assert _device.kde_mode() == mode
    monkeypatch.setattr(_device, "MORTON_MIN_ROWS", 1 << 16)
    monkeypatch.setattr(_device, "MORTON_MIN_PAIRS", 1 << 32)
    rs = np.random.RandomState(40 + D)
    n = K = _device.MORTON_MIN_ROWS + 777
    modes = rs.uniform(-3, 5, size=(6, D))
    c = (modes[rs.randint(0, 6, size=K)] + 0.15 * rs.standard_normal((K, D))).astype(np.float32)
    x = np.concatenate([(modes[rs.randint(0, 6, size=n // 2)] + 0.3 * rs.standard_normal((n // 2, D))),
                        rs.uniform(-4, 6, size=(n - n // 2, D))]).astype(np.float32)
    rs.shuffle(x)
    cd, xd = torch.from_numpy(c).cuda(), torch.from_numpy(x).cuda()
    bw = 0.02
    plain = _device.DeviceMixture.kde(cd, None, K, bw, spatial_order=False)
    ordered = _device.DeviceMixture.kde(cd, None, K, bw)
    assert ordered.spatially_ordered and not getattr(plain, "spatially_ordered", False)
    a = _device.mixture_logpdf(xd, plain).double().cpu().numpy()
    b = _device.mixture_logpdf(xd, ordered).double().cpu().numpy()
    fin = np.isfinite(a)
    assert np.array_equal(np.isfinite(b), fin)
    assert np.max(np.abs(a[fin] - b[fin]) / np.maximum(1.0, np.abs(a[fin]))) < 5e-6  # summation order differs between the two packings
    rows = rs.choice(n, size=300, replace=False)
    ref = O.iso_mixture_log_prob(x[rows].astype(np.float64), c.astype(np.float64), bw)
    ok = np.isfinite(ref)
    assert np.max(np.abs(b[rows][ok] - ref[ok]) / np.maximum(1.0, np.abs(ref[ok]))) < LOGP_TOL
    # fewer points than the threshold: the ordered mixture is evaluated by the ordinary kernel, same values
    few = _device.mixture_logpdf(xd[:5000], ordered).double().cpu().numpy()
    assert np.max(np.abs(few[fin[:5000]] - a[:5000][fin[:5000]]) / np.maximum(1.0, np.abs(a[:5000][fin[:5000]]))) < 5e-6  # summation order differs between the two packings
    # spatial sharding (parallel.spatial_shard): the pieces of the Z-order curve partition the points, are evaluated
    # without the internal sort (points_ordered) and give the ordinary path's values
    from ais_benchmarks_b200 import parallel
    seen = torch.zeros(n, dtype=torch.int32, device="cuda")
    for r in range(3):
        rows_r, idx_r = parallel.spatial_shard(xd, r, 3)
        seen[idx_r] += 1
        assert torch.equal(rows_r, xd[idx_r])
        vals = _device.mixture_logpdf(rows_r, ordered, points_ordered=True).double().cpu().numpy()
        av = a[idx_r.cpu().numpy()]
        fr = np.isfinite(av)
        assert np.max(np.abs(vals[fr] - av[fr]) / np.maximum(1.0, np.abs(av[fr]))) < 5e-6
    assert bool((seen == 1).all())
    # centres chosen by index (with repetitions), as CMixtureSamplingMethod._update_model does
    idx = torch.from_numpy(rs.randint(0, K, size=K)).cuda()
    oi = _device.DeviceMixture.kde(cd, idx, K, bw)
    pi = _device.DeviceMixture.kde(cd, idx, K, bw, spatial_order=False)
    ai = _device.mixture_logpdf(xd, pi).double().cpu().numpy()
    bi = _device.mixture_logpdf(xd, oi).double().cpu().numpy()
    f2 = np.isfinite(ai)
    assert np.max(np.abs(ai[f2] - bi[f2]) / np.maximum(1.0, np.abs(ai[f2]))) < 5e-6  # summation order differs between the two packings


@pytest.mark.parametrize("case", range(3))
def test_tpais_replay_on_device(case):
    """TP-AIS with the reference's captured draws, target densities from the CUDA kernels: same tree and samples,
    weights / NESS within tolerance; haar `prob` (box-mixture kernel, strict bounds) and the leaf mixture model vs the
    reference's values at 400 test points."""
    from test_host_logic import run_tpais_replay
    target = gmm(np.array([[0.0, 0.0], [-0.3, 0.5]]), np.array([[0.01, 0.1], [0.01, 0.03]]), np.array([0.5, 0.5]),
                 [[-1, -1], [1, 1]])
    g, s, samples, weights = run_tpais_replay(case, target)
    assert np.array_equal(s.T.leaves_idx, g["leaves_idx%d" % case])
    assert np.allclose(samples, g["samples%d" % case], atol=1e-12)
    ref_w = g["weights%d" % case]
    big = ref_w > 1e-9 * ref_w.max()
    assert np.max(np.abs(weights[big] - ref_w[big]) / ref_w[big]) < W_TOL
    assert s.get_NESS() == pytest.approx(float(g["ness%d" % case]), rel=W_TOL)
    xt = g["xt%d" % case]
    ref_p = g["prob%d" % case]
    got_p = s.prob(xt)
    assert np.array_equal(got_p == 0, ref_p == 0)
    nz = ref_p > 0
    assert np.max(np.abs(got_p[nz] - ref_p[nz]) / ref_p[nz]) < W_TOL
    with np.errstate(divide="ignore"):
        assert rel_err(s.log_prob(xt), np.log(ref_p)) < 2e-5
    assert rel_err(s.model.log_prob(xt), g["model_logp%d" % case]) < 2e-5
    # end-to-end with device draws: interface contract only (RNG streams are not comparable)
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CTreePyramidSampling
    _device.seed(5)
    t = CTreePyramidSampling({"space_min": np.array([-1.0, -1.0]), "space_max": np.array([1.0, 1.0]), "dims": 2,
                              "method": "simple", "resampling": "leaf", "kernel": str(g["kernel%d" % case]),
                              "ess_target": 0.9, "n_min": 5, "parallel_samples": 32})
    x, w = t.importance_sample(target, 80)
    assert len(x) == len(w) >= 80 and np.all(w > 0) and 0 < t.get_NESS() <= 1
    from ais_benchmarks_b200.metrics import CKLDivergence
    assert np.isfinite(CKLDivergence().compute(p=target, q=t, nsamples=2000))


def test_full_size_config4_kde_kld_row_subset():
    """BASELINE configs[3] at FULL size (1e6 KDE centres x 1e6 evaluation points, D = 8, 16-mode GMM target): the q
    log-density of a random row subset against the oracle, the split/merge identity on the whole output, and the KLD
    against the oracle's KLD formula applied to the device log-densities (the reduction itself)."""
    import bench
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.metrics import CKLDivergence
    from ais_benchmarks_b200.sampling_methods.base import CDeviceKDE
    mu, sig, w, sup, centres, x = bench.kde_workload(scale=1.0)
    target = gmm(mu, sig, w, sup)
    cd, xd = torch.from_numpy(centres).cuda(), torch.from_numpy(x).cuda()
    kde = CDeviceKDE(cd, None, len(centres), 0.02, 8, sup)
    lq = kde.log_prob(xd)
    lp = target.log_prob(xd)
    rows = np.random.RandomState(0).choice(len(x), size=24, replace=False)
    ref_q = O.iso_mixture_log_prob(x[rows].astype(np.float64), centres.astype(np.float64), 0.02, chunk=4)
    assert rel_err(lq.cpu().numpy()[rows], ref_q) < LOGP_TOL, rel_err.last
    assert rel_err(lp.cpu().numpy()[rows], O.gmm_log_prob(x[rows].astype(np.float64), mu, sig, w)) < LOGP_TOL
    h = len(centres) // 2
    a = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd[:h].contiguous(), None, h, 0.02))
    b = _device.mixture_logpdf(xd, _device.DeviceMixture.kde(cd[h:].contiguous(), None, len(centres) - h, 0.02))
    merged = (torch.logaddexp(a.double(), b.double()) + np.log(0.5)).cpu().numpy()
    assert rel_err(lq.cpu().numpy(), merged) < 2e-6, rel_err.last
    val = CKLDivergence().compute_from_samples(target, kde, xd)
    ref = O.kld_from_log_probs(lp.double().cpu().numpy(), lq.double().cpu().numpy())
    assert abs(val - ref) <= 1e-6 * max(1.0, abs(ref))


def test_full_size_config5_dm_weights_row_subset():
    """BASELINE configs[4] at FULL size (1e7 samples x 1024 proposals, D = 32, 64-mode GMM target): log pi, log q and
    log w of a random row subset against the oracle; normaliser / ESS partials against a float64 reduction of the
    device log-weights; normalised weights sum to 1."""
    import bench
    from ais_benchmarks_b200 import _device
    mu, sig, w, sup, pmeans, per = bench.dm_workload(scale=1.0)
    target = gmm(mu, sig, w, sup)
    pm_d = torch.from_numpy(pmeans).cuda().float().contiguous()
    qmix = _device.DeviceMixture.kde(pm_d, None, len(pmeans), 0.05)
    _device.seed(99)
    xs = _device.sample_iso(pm_d, per, 0.05)
    assert xs.shape == (len(pmeans) * per, 32) and xs.shape[0] > 9.9e6
    logw, logp, logq, part = _device.dm_weights(xs, target.device_mixture(), None, qmix, want_logp=True, want_logq=True)
    rows = np.random.RandomState(1).choice(xs.shape[0], size=256, replace=False)
    xr = xs[rows].double().cpu().numpy()
    pm32 = pm_d.double().cpu().numpy()
    assert rel_err(logq[rows].cpu().numpy(), O.iso_mixture_log_prob(xr, pm32, 0.05)) < LOGP_TOL, rel_err.last
    assert rel_err(logp[rows].cpu().numpy(), O.gmm_log_prob(xr, mu, sig, w)) < LOGP_TOL, rel_err.last
    p = part.cpu().numpy()
    lw = logw.double()
    M = float(lw.max())
    assert p[3] == xs.shape[0] and p[0] == pytest.approx(M, abs=1e-6)
    assert p[1] == pytest.approx(float(torch.exp(lw - M).sum()), rel=1e-5)
    assert p[2] == pytest.approx(float(torch.exp(2 * (lw - M)).sum()), rel=1e-5)
    wn = _device.normalize_weights(logw, p)
    assert float(wn.double().sum()) == pytest.approx(1.0, abs=1e-4)


@pytest.mark.parametrize("D,N,per", [(32, 1024, 40), (16, 300, 50), (12, 129, 77), (9, 64, 33), (32, 100, 3)])
def test_tensor_core_proposal_density(D, N, per):
    """tcgen05 screen + exact refinement (csrc/tc.cu) against the oracle on DM-AIS-shaped data (every sample hinted
    with the proposal it was drawn from), against the CUDA-core kernel (must agree to float32 rounding), and with
    deliberately wrong hints (accuracy must hold; only speed may suffer)."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(D + N)
    pm = rs.uniform(-1, 1, size=(N, D))
    pm_d = torch.from_numpy(pm).cuda().float().contiguous()
    mix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
    assert _device.tc_eligible(mix)
    _device.seed(D)
    xs = _device.sample_iso(pm_d, per, 0.05)
    hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
    op = _device.TensorCoreOperand(mix)
    got = _device.tc_logq(xs, op, hint)
    ref32 = _device.mixture_logpdf(xs, mix)
    ref64 = O.iso_mixture_log_prob(xs.double().cpu().numpy(), pm_d.double().cpu().numpy(), 0.05)
    assert rel_err(got.cpu().numpy(), ref64) < LOGP_TOL, rel_err.last
    assert rel_err(got.cpu().numpy(), ref32.double().cpu().numpy()) < 2e-6, rel_err.last
    bad = torch.from_numpy(rs.randint(0, N, size=xs.shape[0]).astype(np.int32)).cuda()
    got_bad = _device.tc_logq(xs, op, bad)
    assert rel_err(got_bad.cpu().numpy(), ref64) < LOGP_TOL, rel_err.last
    # points that belong to no proposal in particular (uniform on the support), arbitrary hints
    xu = torch.from_numpy(rs.uniform(-1, 1, size=(257, D)).astype(np.float32)).cuda()
    hu = torch.zeros(257, dtype=torch.int32, device="cuda")
    refu = O.iso_mixture_log_prob(xu.double().cpu().numpy(), pm_d.double().cpu().numpy(), 0.05)
    assert rel_err(_device.tc_logq(xu, op, hu).cpu().numpy(), refu) < LOGP_TOL, rel_err.last


def test_tensor_core_screen_far_from_origin():
    """Coordinates far from the origin inflate the TF32 screen error (|2 u.v| 2^-10 ~ tens of log2 units here):
    every point widens its refine window by its own bound, so the kernel falls back to exact evaluation instead of
    losing accuracy; the Python dispatcher does not pick the tensor-core path for such a mixture in the first place."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(11)
    D, N, per = 32, 200, 9
    pm = 8.0 + rs.uniform(-1, 1, size=(N, D))
    pm_d = torch.from_numpy(pm).cuda().float().contiguous()
    mix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
    _device.seed(5)
    xs = _device.sample_iso(pm_d, per, 0.05)
    hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
    op = _device.TensorCoreOperand(mix)
    assert op.screen_error > 20 and _device.tc_operand(mix) is None
    ref64 = O.iso_mixture_log_prob(xs.double().cpu().numpy(), pm_d.double().cpu().numpy(), 0.05)
    ref32 = _device.mixture_logpdf(xs, mix).double().cpu().numpy()
    # at |x| ~ 8 the float32 representation of a*mu alone costs 1.4e-5 (CUDA-core kernel included): the tensor-core
    # path must reproduce the CUDA-core result to rounding and stay within that float32 limit of the oracle
    for h in (hint, torch.zeros_like(hint)):
        got = _device.tc_logq(xs, op, h).cpu().numpy()
        assert rel_err(got, ref32) < 2e-6, rel_err.last
        assert rel_err(got, ref64) < 5e-5, rel_err.last
    near = _device.DeviceMixture.kde((pm_d - 8.0).contiguous(), None, N, 0.05)
    assert _device.tc_operand(near) is not None and _device.tc_operand(near).screen_error < 2


def test_dm_weights_hinted_equals_unhinted():
    """aisb_dm_weights_hinted_f32 (tensor-core proposals) vs aisb_dm_weights_f32 (fused CUDA-core kernel): same
    log-weights and partial sums on a 16-D DM-AIS batch; and the sampler routes through it end to end."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
    rs = np.random.RandomState(3)
    D, N, K = 16, 96, 40
    means = rs.uniform(-0.7, 0.7, size=(5, D)); sig = rs.uniform(0.04, 0.05, size=(5, D)); w = rs.uniform(size=5)
    target = gmm(means, sig, w, [np.full(D, -1.0), np.full(D, 1.0)])
    np.random.seed(4)
    _device.seed(4)
    s = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": K, "N": N,
                                  "J": 10, "sigma": 0.05})
    xs = _device.sample_iso(s.proposal_means(), K, 0.05)
    lw_a, pa = s.weigh(xs, target)
    lw_b, pb = s.weigh(xs, target, hint=s._own_proposal())
    assert rel_err(lw_b.cpu().numpy(), lw_a.double().cpu().numpy()) < 5e-6, rel_err.last
    pa, pb = pa.cpu().numpy(), pb.cpu().numpy()
    assert pb[3] == N * K and np.allclose(pa, pb, rtol=1e-5)
    x, wn = s.importance_sample(target, 2 * N * K)
    assert x.shape == (2 * N * K, D) and abs(wn.sum() - 1) < 1e-5 and 0 < s.get_NESS() <= 1


def test_benchmark_fused_evaluation_loop(tmp_path):
    """benchmark/CBenchmark.py: the reference's (target, method) evaluation loop (CBenchmark.py:328-544) on the default
    targets and methods with reduced budgets. Fused (one evaluation set, one q evaluation per iteration, all metrics
    from the same device vectors) and unfused (every metric on its own points, the reference's data flow) runs must
    agree within Monte-Carlo noise, and the results file must parse with the reference's column layout."""
    import os
    import yaml
    from ais_benchmarks_b200.benchmark import CBenchmark
    from ais_benchmarks_b200.metrics import CKLDivergence
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ais_benchmarks_b200", "benchmark")
    bench = yaml.safe_load(open(os.path.join(here, "def_benchmark.yaml")))
    for t in bench["targets"]:
        t["nsamples"], t["nsamples_eval"], t["batch_size"] = 200, 4000, 50
    bfile, cfile, out = tmp_path / "b.yaml", tmp_path / "c.yaml", tmp_path / "res" / "results.txt"
    bfile.write_text(yaml.safe_dump(bench))
    cfile.write_text("nreps: 1\nrseed: 3\nmetrics: [KLD, JSD, EVMSE, T, MEM]\ndebug: {text: false}\n")
    b = CBenchmark()
    b.run(str(bfile), os.path.join(here, "def_methods.yaml"), str(cfile), str(out))
    table = CBenchmark.read_results(str(out))
    assert list(table)[:7] == ["dims", "output_samples", "KLD", "JSD", "EVMSE", "T", "MEM"]
    assert set(table["method"]) == {"m_pmc", "TP_AISr_ESS90", "lais", "dm_ais"} and set(table["target_d"]) == {"gmm", "normal"}
    # M-PMC with sigma = 0.01 on the wide 1-D targets zeroes all its mixture weights (as it does in the reference, whose
    # q is then identically 0: empty inlier sets, NaN expectations); every other (method, target) must be finite
    sane = np.array(table["method"]) != "m_pmc"
    for col in ("KLD", "JSD", "EVMSE", "T", "NESS", "accept_rate"):
        assert np.all(np.isfinite(np.array(table[col])[sane])), col
    assert np.all(np.array(table["JSD"])[sane] >= -1e-6) and np.all(np.array(table["JSD"])[sane] <= 1 + 1e-6)
    assert np.all(np.array(table["T"]) > 0) and max(table["output_samples"]) >= 200

    # fused vs unfused on one (target, method): same estimator, independent evaluation points
    b.load_methods(os.path.join(here, "def_methods.yaml"), b.targets[1].support()[0], b.targets[1].support()[1], 2)
    dm = [m for m in b.methods if m.name == "dm_ais"][0]
    rows = {}
    for fused in (True, False):
        rows[fused] = CBenchmark.evaluate_method(2, b.targets[1], dm, max_samples=400, sampling_eval_samples=40000,
                                                 metrics=["KLD", "JSD", "EVMSE"], rseed=5, n_reps=1, batch_size=100,
                                                 fused=fused)
    assert len(rows[True]) == len(rows[False]) >= 4
    for key in ("KL", "JSD"):
        a, c = rows[True][-1][key], rows[False][-1][key]
        assert abs(a - c) <= 0.1 * max(abs(a), abs(c)) + 0.02, (key, a, c)
    # the fused KLD is exactly the metric on the loop's own evaluation set: the sampler's final model, fresh points
    np.random.seed(9)
    direct = CKLDivergence().compute(p=b.targets[1], q=dm, nsamples=40000)
    assert abs(direct - rows[True][-1]["KL"]) <= 0.1 * abs(direct) + 0.02
    assert dm.device_outputs is False  # restored



@pytest.mark.parametrize("method", ["dm", "lais"])
def test_cuda_graph_adapt_iteration(method):
    """A whole adapt iteration replayed as a CUDA graph (draws from a device-resident Philox counter, proposal packing,
    weight kernel, resampling / MH layer): every iteration's weights match the float64 oracle for the samples it drew
    and the proposal means it started from; replays draw fresh numbers; the eager path gives the same statistics."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS
    rs = np.random.RandomState(4)
    D, N, K = 3, 16, 40
    mu = rs.uniform(-0.6, 0.6, size=(5, D))
    sig = rs.uniform(0.04, 0.06, size=(5, D))
    w = rs.rand(5)
    w /= w.sum()
    target = gmm(mu, sig, w, support=[np.full(D, -1.0), np.full(D, 1.0)])
    params = {"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": K, "N": N, "J": 10,
              "sigma": 0.02, "device_outputs": True, "cuda_graph": True}
    cls = CDeterministicMixtureAIS
    if method == "lais":
        cls = CLayeredAIS
        params.update({"L": 5, "mh_sigma": 0.05})
    np.random.seed(1)
    _device.seed(11)
    s = cls(params)
    seen = []
    for it in range(6):
        means_before = s.proposal_means().double().cpu().numpy().copy()
        x, wn = s.importance_sample(target, (it + 1) * N * K)
        assert s._captured is not None and x.shape[0] == (it + 1) * N * K
        xb = x[-N * K:].double().cpu().numpy()
        lwb = s._logw_dev[-N * K:].double().cpu().numpy()
        ref = O.gmm_log_prob(xb, mu, sig, w) - O.iso_mixture_log_prob(xb, means_before, 0.02)
        fin = np.isfinite(ref) & np.isfinite(lwb)
        assert fin.mean() > 0.9
        assert np.max(np.abs(lwb[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))) < 5e-5
        # proposal-major draws around the means the iteration started from
        d = (xb.reshape(N, K, D) - means_before[:, None, :]) / np.sqrt(0.02)
        assert abs(d.mean()) < 0.15 and abs(d.std() - 1.0) < 0.1
        assert abs(float(wn.double().sum()) - 1.0) < 1e-4
        means_after = s.proposal_means().double().cpu().numpy()
        assert not np.array_equal(means_after, means_before)
        if method == "dm":  # every new mean is one of the iteration's own draws
            x32 = x[-N * K:].cpu().numpy()
            assert all((x32 == m.astype(np.float32)).all(1).any() for m in s.proposal_means().cpu().numpy())
        else:
            assert np.all(means_after >= -1.0) and np.all(means_after <= 1.0)
        seen.append(xb)
    assert not np.array_equal(seen[0], seen[1]) and not np.array_equal(seen[1], seen[2])   # fresh numbers per replay
    assert s.num_proposal_samples >= 6 * N * K and s.get_approx_NESS() > 0
    # reset() puts fresh host-drawn means into the captured buffer; the graph is reused
    cap = s._captured
    s.reset()
    m0 = s.proposal_means().double().cpu().numpy().copy()
    x, _ = s.importance_sample(target, N * K)
    assert s._captured is cap
    d = (x.double().cpu().numpy().reshape(N, K, D) - m0[:, None, :]) / np.sqrt(0.02)
    assert abs(d.std() - 1.0) < 0.1
    # one seed reproduces a run, also through the captured graph (device-resident Philox state, torch generator)
    runs = []
    for _ in range(2):
        np.random.seed(1)
        _device.seed(11)
        r = cls(params)
        xr, wr = r.importance_sample(target, 3 * N * K)
        runs.append((xr.clone(), wr.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    s.reset()
    np.random.seed(1)
    _device.seed(11)
    s.reset()
    xs_, _ = s.importance_sample(target, 3 * N * K)      # the EARLIER capture follows the re-seed as well
    assert torch.equal(xs_, runs[0][0])
    # eager path (cuda_graph False): same adaptation quality
    np.random.seed(1)
    e = cls(dict(params, cuda_graph=False))
    e.importance_sample(target, 6 * N * K)
    assert e._captured is None
    np.random.seed(1)
    g2 = cls(params)
    g2.importance_sample(target, 6 * N * K)
    # both end with proposals concentrated on the target's modes: mean target log-density at the proposal means
    le = O.gmm_log_prob(e.proposal_means().double().cpu().numpy(), mu, sig, w).mean()
    lg = O.gmm_log_prob(g2.proposal_means().double().cpu().numpy(), mu, sig, w).mean()
    assert abs(le - lg) < 2.5, (le, lg)


def test_cuda_graph_edge_cases():
    """Captured iterations at awkward shapes: a single proposal with a single draw, tensor-core-eligible dimensions
    (the captured iteration stays on the CUDA-core weight kernel), a change of target (re-capture), "auto" mode
    switching between the graph (small batches) and eager launches (large batches), a foreign target (never captured)."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
    rs = np.random.RandomState(8)

    def make_target(D, k):
        return (rs.uniform(-0.5, 0.5, size=(k, D)), rs.uniform(0.05, 0.08, size=(k, D)), np.full(k, 1.0 / k))

    def check_last_batch(s, tp, means_before, nk, sigma):
        xb = s._samples_dev[-nk:].double().cpu().numpy()
        lwb = s._logw_dev[-nk:].double().cpu().numpy()
        ref = O.gmm_log_prob(xb, *tp) - O.iso_mixture_log_prob(xb, means_before, sigma)
        fin = np.isfinite(ref) & np.isfinite(lwb)
        assert fin.any() and np.max(np.abs(lwb[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))) < 5e-5

    np.random.seed(2)
    _device.seed(2)
    for D, N, K in [(1, 1, 1), (12, 7, 3), (2, 3, 1000)]:
        tp = make_target(D, 3)
        tgt = gmm(*tp, support=[np.full(D, -1.0), np.full(D, 1.0)])
        s = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": K,
                                      "N": N, "J": 1, "sigma": 0.1, "device_outputs": True, "cuda_graph": True})
        for it in range(3):
            mb = s.proposal_means().double().cpu().numpy().copy()
            s.importance_sample(tgt, (it + 1) * N * K)
            check_last_batch(s, tp, mb, N * K, 0.1)
        cap = s._captured
        assert cap is not None
        tp2 = make_target(D, 2)                       # another target: the iteration is captured again
        tgt2 = gmm(*tp2, support=[np.full(D, -1.0), np.full(D, 1.0)])
        mb = s.proposal_means().double().cpu().numpy().copy()
        s.importance_sample(tgt2, 4 * N * K)
        assert s._captured is not cap
        check_last_batch(s, tp2, mb, N * K, 0.1)
    # "auto": small batches replay a graph, batches above GRAPH_MAX_BATCH launch eagerly
    D = 2
    tp = make_target(D, 3)
    tgt = gmm(*tp, support=[np.full(D, -1.0), np.full(D, 1.0)])
    small = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": 10,
                                      "N": 5, "J": 1, "sigma": 0.1})
    small.importance_sample(tgt, 50)
    assert small._captured is not None
    big = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D,
                                    "K": (1 << 18) // 4 + 1, "N": 4, "J": 1, "sigma": 0.1})
    big.importance_sample(tgt, 10)
    assert big._captured is None and big._n() == 4 * ((1 << 18) // 4 + 1)

    class Foreign:
        def log_prob(self, v):
            return O.gmm_log_prob(v, *tp)
    f = CDeterministicMixtureAIS({"space_min": np.full(D, -1.0), "space_max": np.full(D, 1.0), "dims": D, "K": 10,
                                  "N": 5, "J": 1, "sigma": 0.1, "cuda_graph": True})
    x, w = f.importance_sample(Foreign(), 100)
    assert f._captured is None and x.shape == (100, D) and abs(w.sum() - 1.0) < 1e-5


def test_sharded_samplers_multi_gpu():
    """DM-AIS / LAIS / M-PMC with `sharded: True` under torchrun (scripts/sharded_samplers_check.py): needs >= 2 GPUs
    on the box, skipped otherwise (the host-side exchange logic is covered on CPU by tests/test_parallel_gloo.py)."""
    import os
    import subprocess
    import sys
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(root, "scripts", "sharded_samplers_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "sharded samplers ok on 2 ranks" in res.stdout
