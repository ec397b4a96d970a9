"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run here (the build container), not on the GPU box:  python tests/golden/make_golden.py
Every fixture stores the float64 inputs next to the reference's float64 outputs, so the tests
can feed the same numbers to the oracle (CPU) and to the CUDA path (GPU) without the reference.
RNG streams are never part of a fixture's contract: wherever the reference draws random numbers
inside the code under test (MH noise, KDE resampling), the draws are captured and stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

ref_harness.load()
from ais_benchmarks.distributions import (CGaussianMixtureModel, CMixtureModel, CMultivariateNormal,  # noqa: E402
                                          CMultivariateUniform)
from ais_benchmarks.metrics.divergences import CJSDivergence, CKLDivergence  # noqa: E402
from ais_benchmarks.metrics.statistics import CExpectedValueMSE  # noqa: E402
from ais_benchmarks.sampling_methods import (CDeterministicMixtureAIS, CLayeredAIS, CMixturePMC,  # noqa: E402
                                             CTreePyramidSampling)
from ais_benchmarks.sampling_methods.base import CMixtureSamplingMethod  # noqa: E402
from ais_benchmarks.utils.misc import generateRandomGMM  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024.0))


def random_gmm(seed, dims, k):
    """BASELINE.json configs 2-5 targets: generateRandomGMM (utils/misc.py:40-60), 'sigma' = variances."""
    np.random.seed(seed)
    return generateRandomGMM(np.full(dims, -1.0), np.full(dims, 1.0), k,
                             sigma_min=np.full(dims, 0.04), sigma_max=np.full(dims, 0.05))


def gmm_arrays(g):
    return np.array(g.means, dtype=np.float64), np.array(g.sigmas, dtype=np.float64), np.array(g.weights, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
def make_mvn():
    rs = np.random.RandomState(1)
    out = {}
    cases = []
    # the reference's own test inputs (tests/distributions/test_multivariate_normal_unittest.py:17-48)
    cases.append((np.array([0.0]), np.diag([1.5]), np.array([[0], [0.2], [-2], [-.2]], dtype=np.float64)))
    cases.append((np.zeros(3), np.diag([1.5, 1.5, 1.5]),
                  np.array([[0, 0, 0], [0, .2, 0], [-2, 0, .3], [-.2, .1, .3]], dtype=np.float64)))
    # full covariance, random SPD, D = 4 and D = 8
    for d in (4, 8):
        a = rs.normal(size=(d, d))
        cases.append((rs.normal(size=d), a @ a.T / d + 0.1 * np.eye(d), rs.normal(size=(64, d)) * 1.5))
    # the non-symmetric "covariance" of CMultivariateNormal.py's __main__ demo (:97-99)
    cases.append((np.zeros(2), np.array([[0.1, .2], [.0, .2]]), rs.uniform(-1, 1, size=(64, 2))))
    # default-config Normal target (benchmark/def_benchmark.yaml:26-28)
    cases.append((np.array([0.0]), np.array([[0.2]]), rs.uniform(-5, 5, size=(256, 1))))
    for i, (mean, cov, x) in enumerate(cases):
        dist = CMultivariateNormal({"mean": mean.copy(), "sigma": cov.copy()})
        out["mean%d" % i], out["cov%d" % i], out["x%d" % i] = mean, cov, x
        out["logp%d" % i] = dist.log_prob(x)
        out["prob%d" % i] = dist.prob(x)
        out["support%d" % i] = dist.support()
    out["n"] = np.array(len(cases))
    save("mvn", **out)


def make_gmm():
    rs = np.random.RandomState(2)
    out = {}
    targets = []
    # benchmark/def_benchmark.yaml:4-24 verbatim
    targets.append(CGaussianMixtureModel({"means": [[1.5], [-4.2], [0.01], [7.5]],
                                          "sigmas": [[0.4], [0.05], [0.003], [0.003]],
                                          "weights": [0.4, 0.4, 0.1, 0.1], "support": [[-10], [10]]}))
    targets.append(CGaussianMixtureModel({"means": [[0.0, 0.0], [-0.3, 0.5]], "sigmas": [[0.01, 0.1], [0.01, 0.03]],
                                          "weights": [0.5, 0.5], "support": [[-1, -1], [1, 1]]}))
    targets.append(random_gmm(10, 2, 4))    # config 2 target
    targets.append(random_gmm(11, 8, 16))   # config 3/4 target
    targets.append(random_gmm(12, 32, 64))  # config 5 target
    targets.append(random_gmm(13, 3, 5))    # non power-of-two D
    for i, g in enumerate(targets):
        means, sigmas, weights = gmm_arrays(g)
        sup = g.support()
        n = 2000 if g.dims <= 2 else 384
        x = rs.uniform(sup[0], sup[1], size=(n, g.dims))
        # half of the points near the modes so that |log p| = O(1) cases are covered too
        idx = rs.randint(0, len(means), size=n // 2)
        x[: n // 2] = means[idx] + rs.normal(size=(n // 2, g.dims)) * np.sqrt(sigmas[idx])
        out["means%d" % i], out["sigmas%d" % i], out["weights%d" % i] = means, sigmas, weights
        out["support%d" % i], out["x%d" % i] = sup, x
        out["logp%d" % i] = g.log_prob(x)
        out["prob%d" % i] = g.prob(x)
    out["n"] = np.array(len(targets))
    save("gmm", **out)


def make_mixture():
    """CMixtureModel over full-covariance components with zero / NaN / negative weights (CMixtureModel.py:25-41)."""
    rs = np.random.RandomState(3)
    d, k = 3, 6
    means = rs.uniform(-1, 1, size=(k, d))
    covs = np.empty((k, d, d))
    for j in range(k):
        a = rs.normal(size=(d, d))
        covs[j] = a @ a.T / d * 0.2 + 0.05 * np.eye(d)
    weights = np.array([0.3, 0.0, np.nan, 0.5, -1.0, 0.2])
    models = [CMultivariateNormal({"mean": means[j].copy(), "sigma": covs[j].copy()}) for j in range(k)]
    mix = CMixtureModel({"models": models, "weights": weights.copy(), "dims": d, "support": [[-2] * d, [2] * d]})
    x = rs.uniform(-2, 2, size=(500, d))
    save("mixture", means=means, covs=covs, weights=weights, norm_weights=np.array(mix.weights), x=x,
         logp=mix.log_prob(x), prob=mix.prob(x), logp_single=mix.log_prob(x[0]))


def make_dm_weights():
    """DM-AIS weights exactly as shipped: per-sample loop w = pi(s) / q(s) (sampling_methods/dm_ais.py:58-86)."""
    out = {}
    cfgs = [(2, 4, 64, 4, 0.05, 20), (8, 16, 64, 2, 0.05, 21), (1, None, 5, 10, 0.5, 22)]
    for i, (dims, k, N, K, sigma, seed) in enumerate(cfgs):
        if k is None:  # default config: 1-D Normal target + def_methods.yaml:61-68 DM-AIS
            target = CMultivariateNormal({"mean": np.array([0.0]), "sigma": np.array([[0.2]]),
                                          "support": np.array([[-5.0], [5.0]])})
            out["tmeans%d" % i], out["tsigmas%d" % i], out["tweights%d" % i] = \
                np.array([[0.0]]), np.array([[0.2]]), np.array([1.0])
        else:
            target = random_gmm(seed, dims, k)
            out["tmeans%d" % i], out["tsigmas%d" % i], out["tweights%d" % i] = gmm_arrays(target)
        np.random.seed(seed + 100)
        sup = target.support()
        s = CDeterministicMixtureAIS({"space_min": sup[0], "space_max": sup[1], "dims": dims,
                                      "K": K, "N": N, "J": 1000, "sigma": sigma})
        out["pmeans%d" % i] = np.array([p.loc for p in s.proposals])
        samples, norm_w = s.importance_sample(target, N * K, timeout=600)  # exactly one iteration
        out["x%d" % i] = np.array(samples)
        out["w%d" % i] = np.array(s.weights)
        out["norm_w%d" % i] = np.array(norm_w)
        out["ness%d" % i] = np.array(s.get_NESS())
        out["sigma%d" % i] = np.array(sigma)
        # proposal mixture density at the samples with the PRE-adaptation proposals
        pre = CMixtureModel({"models": [CMultivariateNormal({"mean": m, "sigma": np.eye(dims) * sigma})
                                        for m in out["pmeans%d" % i]],
                             "weights": np.full(N, 1.0 / N), "dims": dims, "support": [sup[0], sup[1]]})
        out["logq%d" % i] = pre.log_prob(np.array(samples))
        out["logp%d" % i] = target.log_prob(np.array(samples))
    out["n"] = np.array(len(cfgs))
    save("dm_weights", **out)


class _KdeHolder(CMixtureSamplingMethod):
    """Smallest concrete user of the reference's KDE builder (sampling_methods/base.py:135-179)."""

    def __init__(self, params):
        super().__init__(params)
        self.bw = params["kde_bw"]

    def importance_sample(self, target_d, n_samples, timeout):
        raise NotImplementedError


def make_kde_kld():
    out = {}
    cfgs = [(2, 4, 400, 100, 0.02, 2000, 30), (8, 16, 3000, 1000, 0.02, 1024, 31), (1, 3, 200, 100, 0.01, 2000, 32)]
    for i, (dims, k, n_samples, n_kde, bw, n_eval, seed) in enumerate(cfgs):
        target = random_gmm(seed, dims, k)
        sup = target.support()
        np.random.seed(seed + 100)
        holder = _KdeHolder({"space_min": sup[0], "space_max": sup[1], "dims": dims, "n_samples_kde": n_kde,
                             "kde_bw": bw})
        means, sigmas, weights = gmm_arrays(target)
        comp = np.random.randint(0, k, size=n_samples)
        holder.samples = means[comp] + np.random.normal(size=(n_samples, dims)) * np.sqrt(sigmas[comp])
        holder._update_model()
        centres = np.array([m.loc for m in holder.model.models])
        x = np.random.uniform(sup[0], sup[1], size=(n_eval, dims))
        lq = holder.log_prob(x)
        lp = target.log_prob(x)
        metric = CKLDivergence()
        kld = metric.compute_from_samples(target, holder, x)
        out["tmeans%d" % i], out["tsigmas%d" % i], out["tweights%d" % i] = means, sigmas, weights
        out["centres%d" % i], out["bw%d" % i], out["x%d" % i] = centres, np.array(bw), x
        out["lp%d" % i], out["lq%d" % i], out["kld%d" % i] = lp, lq, np.array(kld)
        out["terms%d" % i] = CKLDivergence.compute_from_log_probs(lp.copy(), lq.copy())
    out["n"] = np.array(len(cfgs))
    # closed-form Gaussian cases of tests/metrics/test_metric_kldivergence.py:77-135 on seeded uniform points
    kat = [([0.], [0.01], [[1.]], [[1.]]), ([0.], [2.], [[1.]], [[1.]]), ([0.], [5.], [[1.]], [[.1]]),
           ([0.], [2.], [[2.]], [[2.]]), ([0., 0., 0.], [5., 5., 5.], np.eye(3).tolist(), (np.eye(3) * .5).tolist())]
    for j, (m0, m1, s0, s1) in enumerate(kat):
        p = CMultivariateNormal({"mean": np.array(m0), "sigma": np.array(s0)})
        q = CMultivariateNormal({"mean": np.array(m1), "sigma": np.array(s1)})
        np.random.seed(40 + j)
        out["kat_m0_%d" % j], out["kat_m1_%d" % j] = np.array(m0), np.array(m1)
        out["kat_s0_%d" % j], out["kat_s1_%d" % j] = np.array(s0), np.array(s1)
        out["kat_seed_%d" % j] = np.array(40 + j)
        out["kat_kld_%d" % j] = np.array(CKLDivergence().compute(p=p, q=q, nsamples=100000))
    out["n_kat"] = np.array(len(kat))
    # edge cases of compute_from_log_probs: -inf entries on either side, empty inlier set
    lp = np.array([-1.0, -np.inf, -2.0, -0.5, -3.0])
    lq = np.array([-1.5, -1.0, -np.inf, -0.7, -2.0])
    out["edge_lp"], out["edge_lq"] = lp, lq
    out["edge_kld"] = np.array(np.sum(CKLDivergence.compute_from_log_probs(lp.copy(), lq.copy())))
    out["empty_kld"] = np.array(np.sum(CKLDivergence.compute_from_log_probs(np.full(4, -np.inf), np.zeros(4))))
    save("kde_kld", **out)


class _FixedProbs:
    """Stand-in distribution whose prob() returns a fixed vector (edge cases of the linear-domain metrics)."""

    def __init__(self, probs):
        self.probs = np.asarray(probs, dtype=np.float64)

    def prob(self, samples):
        return self.probs.copy()

    log_prob = prob


def make_metrics():
    """JSD (metrics/divergences.py:156-195) and EVMSE (metrics/statistics.py:29-48) of the reference on the KDE
    fixtures of kde_kld.npz (same p, q, evaluation points), plus edge cases with zero / NaN probabilities."""
    g = np.load(os.path.join(OUT, "kde_kld.npz"))
    out = {}
    for i in range(int(g["n"])):
        means, sigmas, weights = g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i]
        dims = means.shape[1]
        sup = [np.full(dims, -1.0), np.full(dims, 1.0)]
        p = CGaussianMixtureModel({"means": means, "sigmas": sigmas, "weights": weights, "support": sup})
        centres = g["centres%d" % i]
        q = CGaussianMixtureModel({"means": centres, "sigmas": np.full(centres.shape, float(g["bw%d" % i])),
                                   "support": sup})
        x = g["x%d" % i]
        assert np.allclose(q.log_prob(x), g["lq%d" % i], rtol=1e-12, atol=1e-12)
        out["jsd%d" % i] = np.array(CJSDivergence().compute_from_samples(p, q, x.copy()))
        out["evmse%d" % i] = np.array(CExpectedValueMSE().compute_from_samples(p, q, x.copy()))
        pp, qq = p.prob(x), q.prob(x)
        out["ev_p%d" % i] = np.sum(x * (pp / pp.sum()).reshape(-1, 1), axis=0)
        out["ev_q%d" % i] = np.sum(x * (qq / qq.sum()).reshape(-1, 1), axis=0)
    # closed-form-ish Gaussian pair: JSD of two unit normals one sigma apart, on seeded uniform points
    pn = CMultivariateNormal({"mean": np.array([0.0]), "sigma": np.array([[1.0]])})
    qn = CMultivariateNormal({"mean": np.array([1.0]), "sigma": np.array([[1.0]])})
    np.random.seed(70)
    xn = np.random.uniform(pn.support()[0], pn.support()[1], size=(20000, 1))
    out["norm_x"] = xn
    out["norm_jsd"] = np.array(CJSDivergence().compute_from_samples(pn, qn, xn.copy()))
    out["norm_evmse"] = np.array(CExpectedValueMSE().compute_from_samples(pn, qn, xn.copy()))
    # edge cases: zero probabilities on either side, NaN, a value that underflows after normalisation
    ep = np.array([0.2, 0.0, 0.3, 0.1, np.nan, 0.4, 1e-300, 0.05])
    eq = np.array([0.1, 0.2, 0.0, 0.3, 0.2, 0.4, 0.3, np.nan])
    out["edge_p"], out["edge_q"] = ep, eq
    with np.errstate(all="ignore"):
        out["edge_jsd"] = np.array(CJSDivergence().compute_from_samples(_FixedProbs(ep), _FixedProbs(eq), np.zeros((8, 1))))
    ex = np.arange(12, dtype=np.float64).reshape(6, 2) / 7.0
    fp, fq = np.array([0.2, 0.0, 0.3, 0.1, 0.9, 0.4]), np.array([0.1, 0.2, 0.0, 0.3, 0.2, 0.4])
    out["edge_x"], out["edge_fp"], out["edge_fq"] = ex, fp, fq
    out["edge_evmse"] = np.array(CExpectedValueMSE().compute_from_samples(_FixedProbs(fp), _FixedProbs(fq), ex.copy()))
    out["n"] = np.array(int(g["n"]))
    save("metrics", **out)


def make_mpmc():
    """Two M-PMC iterations (sampling_methods/m_pmc.py:80-118): the first from isotropic proposals, the second from the
    adapted ones (full covariance when the 10*I failsafe does not fire)."""
    out = {}
    cfgs = [(2, 0.3, 6, 40, 0.001, 50), (3, 1.0, 8, 60, 0.01, 51)]
    for i, (dims, half, N, K, sigma, seed) in enumerate(cfgs):
        np.random.seed(seed)
        target = generateRandomGMM(np.full(dims, -half), np.full(dims, half), 3,
                                   sigma_min=np.full(dims, 0.002), sigma_max=np.full(dims, 0.004))
        out["tmeans%d" % i], out["tsigmas%d" % i], out["tweights%d" % i] = gmm_arrays(target)
        np.random.seed(seed + 100)
        s = CMixturePMC({"space_min": np.full(dims, -half), "space_max": np.full(dims, half), "dims": dims,
                         "K": K, "N": N, "J": 1000, "sigma": sigma})
        for it in range(2):
            tag = "%d_%d" % (i, it)
            out["means" + tag] = np.array([p.loc for p in s.proposals])
            out["covs" + tag] = np.array([p.scale for p in s.proposals])
            out["alpha" + tag] = np.array(s.wproposals)
            out["mixw" + tag] = np.array(s.model.weights)
            n_before = len(s.samples)
            s.importance_sample(target, n_before + K, timeout=600)
            out["x" + tag] = np.array(s.samples[n_before:])
            out["w" + tag] = np.array(s.weights[n_before:])
            out["new_means" + tag] = np.array([p.loc for p in s.proposals])
            out["new_covs" + tag] = np.array([p.scale for p in s.proposals])
            out["new_alpha" + tag] = np.array(s.wproposals)
        out["sigma%d" % i] = np.array(sigma)
    out["n"] = np.array(len(cfgs))
    save("mpmc", **out)


def _f32(a):
    """Round captured draws to float32-representable float64: the reference then computes from exactly the numbers the
    float32 device path receives, and the fixture stores them in half the space (draws are inputs, not results)."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def make_config3():
    """BASELINE configs[2] SHAPE (8-D 16-mode GMM target, 256 proposals) driven through the reference's own adapt loops:

    * M-PMC (sampling_methods/m_pmc.py:53-118), two consecutive iterations, in two geometries:
        "std"  support [-1, 1]^8, sigma 0.05, K = 10 000 draws per iteration. The raw scatter of quirk Q3 exceeds 10 in
               both iterations, so both updates end in the 10*I failsafe (m_pmc.py:71-72).
        "tiny" support [-0.01, 0.01]^8, sigma 1e-6, K = 4 000. Iteration 1 keeps the raw scatter below 10: a genuine
               FULL covariance for all 256 proposals (no failsafe); iteration 2 draws from those full-covariance
               proposals and weighs against them, and its update hits the failsafe.
      The mixture draws are captured (rounded to float32) by wrapping self.sample; everything else is reference code.
    * LAIS (sampling_methods/layered_ais.py:54-104): DM weights of one 256 x 39 batch against the 256 proposals
      (:90-94) and the 256-chain x 10-step MH layer with captured noise (:54-74).
    """
    out = {}
    dims, N = 8, 256
    geoms = [("std", 1.0, 0.05, 10000, (0.04, 0.05), 80), ("tiny", 0.01, 1e-6, 4000, (4e-6, 5e-6), 81)]
    for name, half, sigma, K, (smin, smax), seed in geoms:
        np.random.seed(seed)
        target = generateRandomGMM(np.full(dims, -half), np.full(dims, half), 16,
                                   sigma_min=np.full(dims, smin), sigma_max=np.full(dims, smax))
        out["tmeans_" + name], out["tsigmas_" + name], out["tweights_" + name] = gmm_arrays(target)
        np.random.seed(seed + 100)
        s = CMixturePMC({"space_min": np.full(dims, -half), "space_max": np.full(dims, half), "dims": dims,
                         "K": K, "N": N, "J": 1000, "sigma": sigma})
        real_sample = s.sample
        s.sample = lambda n, _f=real_sample: _f32(_f(n))
        for it in range(2):
            tag = "_%s_%d" % (name, it)
            out["means" + tag] = np.array([p.loc for p in s.proposals])
            out["covs" + tag] = np.array([p.scale for p in s.proposals])
            out["alpha" + tag] = np.array(s.wproposals)
            out["mixw" + tag] = np.array(s.model.weights)
            n_before = len(s.samples)
            s.importance_sample(target, n_before + K, timeout=3600)
            x = np.array(s.samples[n_before:])
            assert len(x) == K and np.array_equal(x, _f32(x))
            out["x" + tag] = x.astype(np.float32)
            out["w" + tag] = np.array(s.weights[n_before:])
            out["new_means" + tag] = np.array([p.loc for p in s.proposals])
            out["new_covs" + tag] = np.array([p.scale for p in s.proposals])
            out["new_alpha" + tag] = np.array(s.wproposals)
            failsafe = np.all(out["new_covs" + tag] == np.diag(np.ones(dims) * 10)[None], axis=(1, 2))
            out["failsafe" + tag] = failsafe
            print("  m-pmc %-4s it %d: failsafe on %d / %d proposals, sum w = %.6f" % (name, it, failsafe.sum(), N,
                                                                                     out["w" + tag].sum()))
        out["sigma_" + name], out["half_" + name] = np.array(sigma), np.array(half)
    # ---- LAIS at the same shape ----
    seed = 90
    target = random_gmm(seed, dims, 16)
    sup = target.support()
    K, L, sigma, mh_sigma = 39, 10, 0.05, 0.5
    np.random.seed(seed + 100)
    s = CLayeredAIS({"space_min": sup[0], "space_max": sup[1], "dims": dims, "K": K, "N": N, "J": 1000, "L": L,
                     "sigma": sigma, "mh_sigma": mh_sigma})
    out["l_tmeans"], out["l_tsigmas"], out["l_tweights"] = gmm_arrays(target)
    out["l_before"] = np.array([p.loc for p in s.proposals])
    for p in s.proposals:                                    # capture the proposal draws (dm-style sampling, :82-88)
        p.sample = (lambda n_samples=1, _f=p.sample: _f32(_f(n_samples)))
    rs = np.random.RandomState(seed + 200)
    delta = _f32(rs.normal(size=(L, N, dims)) * np.sqrt(mh_sigma))
    u = _f32(rs.uniform(size=(L, N)))
    d_iter = iter([delta[l, n] for n in range(N) for l in range(L)])
    u_iter = iter([u[l, n] for n in range(N) for l in range(L)])
    s.mhproposal.sample = lambda n_samples=1: np.array([next(d_iter)])
    real_rand = np.random.rand
    np.random.rand = lambda *a: next(u_iter)
    try:
        x, w = s.importance_sample(target, N * K, timeout=3600)          # exactly one iteration
    finally:
        np.random.rand = real_rand
    assert len(x) == N * K and np.array_equal(x, _f32(x))
    out["l_x"], out["l_w"] = np.asarray(x, dtype=np.float32), np.asarray(w)
    out["l_ness"] = np.array(s.get_NESS())
    out["l_after"] = np.array([p.loc for p in s.proposals])
    out["l_delta"], out["l_u"] = delta.astype(np.float32), u.astype(np.float32)
    out["l_lo"], out["l_hi"] = sup[0], sup[1]
    out["l_cfg"] = np.array([K, L, sigma, mh_sigma])
    out["l_pi_evals"] = np.array(s.num_target_evals)
    _config3_logs(out)
    save("config3", **out)


def _config3_logs(out, rows=1000):
    """Reference log densities behind every M-PMC batch of make_config3, for the first `rows` draws: log pi(x) and
    log q(x) under the PRE-update proposal mixture (CMixtureModel.log_prob over CMultivariateNormal components with the
    stored means / covariances / sanitised weights). They stay finite where the linear-domain weights of the reference
    underflow to 0 / NaN ("tiny" iteration 1: proposals ~3e4 times wider than the target), so they are what pins the
    full-covariance 256-component density at this shape."""
    for name in ("std", "tiny"):
        half = float(out["half_" + name])
        sup = [np.full(8, -half), np.full(8, half)]
        target = CGaussianMixtureModel({"means": out["tmeans_" + name], "sigmas": out["tsigmas_" + name],
                                        "weights": out["tweights_" + name], "support": sup})
        for it in range(2):
            tag = "_%s_%d" % (name, it)
            x = np.asarray(out["x" + tag][:rows], dtype=np.float64)
            models = [CMultivariateNormal({"mean": m.copy(), "sigma": c.copy()})
                      for m, c in zip(out["means" + tag], out["covs" + tag])]
            q = CMixtureModel({"models": models, "weights": np.array(out["mixw" + tag], dtype=np.float64),
                               "dims": 8, "support": sup})
            out["lp" + tag] = target.log_prob(x)
            out["lq" + tag] = q.log_prob(x)


def make_hetero():
    """A CMixtureModel over a HETEROGENEOUS component list (CMixtureModel.py:53-56 takes any models[] with log_prob):
    two Gaussians (one full covariance), two boxes and a zero-weight component, 2-D."""
    rs = np.random.RandomState(123)
    comps = [CMultivariateNormal({"mean": np.array([0.3, -0.2]), "sigma": np.diag([0.05, 0.02])}),
             CMultivariateUniform({"center": np.array([-0.4, 0.1]), "radius": np.array([0.3, 0.2])}),
             CMultivariateNormal({"mean": np.array([-0.5, 0.5]), "sigma": np.array([[0.04, 0.01], [0.01, 0.03]])}),
             CMultivariateUniform({"center": np.array([0.5, 0.5]), "radius": np.array([0.25])}),
             CMultivariateNormal({"mean": np.array([0.9, 0.9]), "sigma": np.diag([0.01, 0.01])})]
    w = np.array([0.3, 0.2, 0.25, 0.25, 0.0])
    mix = CMixtureModel({"models": comps, "weights": w.copy(), "dims": 2, "support": [[-1, -1], [1, 1]]})
    x = rs.uniform(-1, 1, size=(400, 2))
    save("hetero", x=x, w=w, logp=mix.log_prob(x), prob=mix.prob(x))


def make_lais():
    """LAIS upper layer (sampling_methods/layered_ais.py:54-74) with the random draws captured."""
    out = {}
    cfgs = [(2, 4, 8, 10, 0.05, 0.5, 60), (8, 16, 16, 5, 0.05, 0.05, 61)]
    for i, (dims, k, N, L, sigma, mh_sigma, seed) in enumerate(cfgs):
        target = random_gmm(seed, dims, k)
        sup = target.support()
        np.random.seed(seed + 100)
        s = CLayeredAIS({"space_min": sup[0], "space_max": sup[1], "dims": dims, "K": 2, "N": N, "J": 1000, "L": L,
                         "sigma": sigma, "mh_sigma": mh_sigma})
        before = np.array([p.loc for p in s.proposals])
        rs = np.random.RandomState(seed + 200)
        delta = rs.normal(size=(L, N, dims)) * np.sqrt(mh_sigma)
        u = rs.uniform(size=(L, N))
        # reference order: for each proposal n, L steps; each step draws delta then u (layered_ais.py:55-62)
        d_iter = iter([delta[l, n] for n in range(N) for l in range(L)])
        u_iter = iter([u[l, n] for n in range(N) for l in range(L)])
        s.mhproposal.sample = lambda n_samples=1: np.array([next(d_iter)])
        real_rand = np.random.rand
        np.random.rand = lambda *a: next(u_iter)
        try:
            s.resample(target)
        finally:
            np.random.rand = real_rand
        out["tmeans%d" % i], out["tsigmas%d" % i], out["tweights%d" % i] = gmm_arrays(target)
        out["before%d" % i], out["after%d" % i] = before, np.array([p.loc for p in s.proposals])
        out["delta%d" % i], out["u%d" % i] = delta, u
        out["lo%d" % i], out["hi%d" % i] = sup[0], sup[1]
        out["pi_evals%d" % i] = np.array(s.num_target_evals)
    out["n"] = np.array(len(cfgs))
    save("lais", **out)


def make_uniform():
    """CMultivariateUniform known answers (tests/distributions/test_multivariate_uniform_unittest.py:16-60) and a
    mixture of uniforms as TP-AIS builds it (sampling_methods/tree_pyramid.py:710-736)."""
    rs = np.random.RandomState(7)
    out = {}
    cases = [(np.array([0.0]), np.array([0.5])), (np.array([0.0, 0.1]), np.array([0.5, 0.2])),
             (np.array([0.0, 0.0, 0.0]), np.array([0.5]))]
    for i, (c, r) in enumerate(cases):
        d = CMultivariateUniform({"center": c.copy(), "radius": r.copy()})
        x = rs.uniform(-1, 1, size=(200, len(c)))
        x[0] = c + np.broadcast_to(r, c.shape)   # upper edge is inside (x <= max)
        x[1] = c - np.broadcast_to(r, c.shape)   # lower edge is outside (min < x)
        out["c%d" % i], out["r%d" % i], out["x%d" % i] = c, r, x
        out["logp%d" % i], out["prob%d" % i] = d.log_prob(x), d.prob(x)
    out["n"] = np.array(len(cases))
    k, dims = 9, 2
    centers = rs.uniform(-1, 1, size=(k, dims))
    radii = rs.uniform(0.1, 0.6, size=(k, 1))
    w = rs.uniform(size=k)
    w[3] = 0.0
    models = [CMultivariateUniform({"center": centers[j].copy(), "radius": radii[j].copy()}) for j in range(k)]
    mix = CMixtureModel({"models": models, "weights": w.copy(), "dims": dims, "support": [[-2, -2], [2, 2]]})
    x = rs.uniform(-1.5, 1.5, size=(600, dims))
    out["mix_centers"], out["mix_radii"], out["mix_w"], out["mix_x"] = centers, radii, w, x
    out["mix_norm_w"] = np.array(mix.weights)
    out["mix_logp"] = mix.log_prob(x)
    save("uniform", **out)


def make_tpais():
    """TP-AIS "simple" runs (sampling_methods/tree_pyramid.py:421-457) with every random draw captured in prototype units
    (uniform: (x - centre) / (2 radius); normal: (x - mean) / sqrt(var)), so that the run can be replayed exactly."""
    out = {}
    cfgs = [("haar", 0.9, 5, 32, 120, 70), ("normal", 0.9, 5, 32, 100, 71), ("haar", 2.0, 5, 8, 60, 72)]
    for i, (kernel, ess_target, n_min, par, n_samples, seed) in enumerate(cfgs):
        target = CGaussianMixtureModel({"means": [[0.0, 0.0], [-0.3, 0.5]], "sigmas": [[0.01, 0.1], [0.01, 0.03]],
                                        "weights": [0.5, 0.5], "support": [[-1, -1], [1, 1]]})
        queue = []
        real_u, real_n = CMultivariateUniform.sample, CMultivariateNormal.sample

        def rec_u(self, n_samples=1):
            res = real_u(self, n_samples)
            queue.append((res - self._loc) / (2 * self._scale))
            return res

        def rec_n(self, n_samples=1):
            res = real_n(self, n_samples)
            queue.append((res - self._loc.reshape(1, -1)) / np.sqrt(np.diag(self._scale)).reshape(1, -1))
            return res

        CMultivariateUniform.sample, CMultivariateNormal.sample = rec_u, rec_n
        try:
            np.random.seed(seed)
            s = CTreePyramidSampling({"space_min": np.array([-1.0, -1.0]), "space_max": np.array([1.0, 1.0]), "dims": 2,
                                      "method": "simple", "resampling": "leaf", "kernel": kernel,
                                      "ess_target": ess_target, "n_min": n_min, "parallel_samples": par})
            samples, weights = s.importance_sample(target, n_samples, timeout=600)
        finally:
            CMultivariateUniform.sample, CMultivariateNormal.sample = real_u, real_n
        rs = np.random.RandomState(seed + 1)
        xt = rs.uniform(-1, 1, size=(400, 2))
        out["kernel%d" % i] = np.array(kernel)
        out["cfg%d" % i] = np.array([ess_target, n_min, par, n_samples], dtype=np.float64)
        out["draws%d" % i] = np.concatenate(queue)
        out["samples%d" % i], out["weights%d" % i] = np.array(samples), np.array(weights)
        out["centers%d" % i], out["radii%d" % i] = np.array(s.T.centers), np.array(s.T.radii).reshape(-1)
        out["tweights%d" % i], out["leaves_idx%d" % i] = np.array(s.T.weights), np.array(s.T.leaves_idx)
        out["ness%d" % i] = np.array(s.get_NESS())
        out["counters%d" % i] = np.array([s.num_proposal_samples, s.num_proposal_evals, s.num_target_evals])
        out["xt%d" % i] = xt
        out["prob%d" % i] = s.prob(xt)
        out["model_logp%d" % i] = s.model.log_prob(xt)
        out["acc%d" % i] = np.array(s.get_acceptance_rate())
    out["n"] = np.array(len(cfgs))
    save("tpais", **out)


if __name__ == "__main__":
    np.seterr(all="ignore")
    if sys.argv[1:] == ["config3_logs"]:       # re-derive the log-density arrays of an existing config3.npz
        with np.load(os.path.join(OUT, "config3.npz")) as f:
            out = {k: f[k] for k in f.files}
        _config3_logs(out)
        save("config3", **out)
        sys.exit(0)
    if len(sys.argv) > 1:                      # regenerate selected fixtures only: make_golden.py config3 mpmc ...
        for name in sys.argv[1:]:
            globals()["make_" + name]()
        sys.exit(0)
    make_mvn()
    make_gmm()
    make_mixture()
    make_dm_weights()
    make_kde_kld()
    make_metrics()
    make_mpmc()
    make_config3()
    make_hetero()
    make_lais()
    make_uniform()
    make_tpais()
