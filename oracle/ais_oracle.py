"""CPU oracle: a numpy float64 restatement of the ais-benchmarks hot path.

TEST INFRASTRUCTURE ONLY. Nothing in the product package (ais_benchmarks_b200/) imports this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
do, and there only as the checker or as the timed CPU baseline.

Parity status: PINNED. Every function below is checked in tests/test_oracle_golden.py against
(a) fixtures produced by running the UNMODIFIED reference (/root/reference, imported with
matplotlib/guppy stubbed) on seeded inputs -- tests/golden/*.npz, generator
tests/golden/make_golden.py -- and (b) the known answers held by the reference's own tests
(scipy.stats.multivariate_normal.pdf values of tests/distributions/test_multivariate_normal_unittest.py:17-48,
the 1/prod(2r) cases of test_multivariate_uniform_unittest.py:16-60 and the closed-form Gaussian KL values of
tests/metrics/test_metric_kldivergence.py:77-135).

Each function names the reference lines it restates (paths relative to the reference root
ais_benchmarks/). Everything is float64 and batch-first (N, D) -> (N,), like the reference.
"""
import numpy as np
from scipy.special import logsumexp

LOG_2PI = np.log(2.0 * np.pi)


# ----------------------------------------------------------------------------------------------
# distributions
# ----------------------------------------------------------------------------------------------
def check_shape(x, dims):
    """distributions/distributions.py:128-147."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        return x.reshape(1, dims)
    if x.ndim == 2 and x.shape[1] == dims:
        return x
    raise ValueError("Shape of x does not match dims = %d: %s" % (dims, str(x.shape)))


def mvn_log_prob(x, mean, cov):
    """distributions/parametric/CMultivariateNormal.py:50,55-62,67-71:
    term1 = -D/2 log 2pi, term2 = -1/2 log det(cov), term3 = -1/2 (mu - x)^T inv(cov) (mu - x)."""
    mean = np.asarray(mean, dtype=np.float64).reshape(-1)
    cov = np.asarray(cov, dtype=np.float64)
    dims = mean.shape[0]
    x = check_shape(x, dims)
    det = np.linalg.det(cov)
    assert det > 0
    inv_cov = np.linalg.inv(cov)
    diff = mean.reshape(1, dims) - x
    term3 = -0.5 * np.einsum("ni,ij,nj->n", diff, inv_cov, diff)
    return -0.5 * dims * LOG_2PI - 0.5 * np.log(det) + term3


def mvn_prob(x, mean, cov):
    """CMultivariateNormal.py:73-74."""
    return np.exp(mvn_log_prob(x, mean, cov))


def mvn_default_support(mean, cov):
    """CMultivariateNormal.py:37-40: mean -/+ 6 sqrt(diag cov)."""
    mean = np.asarray(mean, dtype=np.float64)
    sd = np.sqrt(np.diag(np.asarray(cov, dtype=np.float64)))
    return np.array([mean - 6.0 * sd, mean + 6.0 * sd])


def sanitize_weights(weights):
    """distributions/mixture/CMixtureModel.py:25-41: NaN / <= 0 -> 0; divide by the sum if it is > 0."""
    w = np.array(weights, dtype=np.float64)
    w[np.logical_or(np.isnan(w), w <= 0)] = 0
    if w.sum() > 0:
        w = w / w.sum()
    return w


def mixture_log_prob(x, means, covs, weights):
    """CMixtureModel.py:43-57: per-component log densities stacked (K, N), then logsumexp(b=weights, axis=0).
    means (K, D); covs (K, D, D); weights (K,) raw (sanitised here like the constructor does)."""
    means = np.asarray(means, dtype=np.float64)
    K, dims = means.shape
    x = check_shape(x, dims)
    w = sanitize_weights(weights)
    ll = np.empty((K, x.shape[0]))
    for k in range(K):
        ll[k] = mvn_log_prob(x, means[k], covs[k])
    with np.errstate(divide="ignore"):
        return logsumexp(ll, b=w.reshape(-1, 1), axis=0)


def mixture_prob(x, means, covs, weights):
    """CMixtureModel.py:59-77: sum_k w_k exp(ll_k), linear domain (underflows to 0 like the reference)."""
    means = np.asarray(means, dtype=np.float64)
    K, dims = means.shape
    x = check_shape(x, dims)
    w = sanitize_weights(weights)
    acc = np.zeros(x.shape[0])
    for k in range(K):
        acc += np.exp(mvn_log_prob(x, means[k], covs[k])) * w[k]
    return acc


def diag_covs(sigmas):
    """distributions/mixture/CGaussianMixtureModel.py:29-30: variances go on the covariance diagonal."""
    sigmas = np.asarray(sigmas, dtype=np.float64)
    return np.stack([np.diag(s) for s in sigmas])


def gmm_log_prob(x, means, sigmas, weights=None):
    """CGaussianMixtureModel.py:8-41: diagonal GMM, default equal weights (:35)."""
    means = np.asarray(means, dtype=np.float64)
    if weights is None:
        weights = np.full(len(means), 1.0 / len(means))
    return mixture_log_prob(x, means, diag_covs(sigmas), weights)


def gmm_prob(x, means, sigmas, weights=None):
    means = np.asarray(means, dtype=np.float64)
    if weights is None:
        weights = np.full(len(means), 1.0 / len(means))
    return mixture_prob(x, means, diag_covs(sigmas), weights)


def iso_mixture_log_prob(x, centres, var, weights=None, chunk=4096):
    """Equal-covariance (var * I) mixture, the shape of both the KDE of sampling_methods/base.py:162-179
    (centres = resampled samples, var = bw, weights 1/n_kde) and the DM-AIS proposal mixture of dm_ais.py:36-45.
    Same arithmetic as mixture_log_prob, evaluated in row chunks so that large K stays in memory."""
    centres = np.asarray(centres, dtype=np.float64)
    K, dims = centres.shape
    x = check_shape(x, dims)
    w = sanitize_weights(np.full(K, 1.0 / K) if weights is None else weights)
    const = -0.5 * dims * LOG_2PI - 0.5 * dims * np.log(var)
    out = np.empty(x.shape[0])
    with np.errstate(divide="ignore"):
        for s in range(0, x.shape[0], chunk):
            xs = x[s:s + chunk]
            d2 = ((xs[:, None, :] - centres[None, :, :]) ** 2).sum(-1)
            out[s:s + chunk] = logsumexp(const - 0.5 * d2 / var, b=w.reshape(1, -1), axis=1)
    return out


def uniform_log_prob(x, center, radius):
    """distributions/parametric/CMultivariateUniform.py:38-47: inlier iff min < x <= max in every dimension."""
    center = np.asarray(center, dtype=np.float64).reshape(-1)
    dims = center.shape[0]
    radius = np.broadcast_to(np.asarray(radius, dtype=np.float64).reshape(-1), (dims,))
    x = check_shape(x, dims)
    inl = np.all(np.logical_and(center - radius < x, x <= center + radius), axis=1)
    res = np.full(x.shape[0], -np.sum(np.log(2.0 * radius)))
    res[~inl] = -np.inf
    return res


def box_mixture_log_prob(x, lo, hi, weights, open_upper):
    """log sum_k w_k [x in box_k] / vol_k. open_upper=True is the TP-AIS haar test lo < x < hi
    (sampling_methods/tree_pyramid.py:593-602); False is CMultivariateUniform's lo < x <= hi inside a
    CMixtureModel (CMixtureModel.py:43-57)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    x = check_shape(x, lo.shape[1])
    w = np.asarray(weights, dtype=np.float64)
    dens = np.zeros(x.shape[0])
    for k in range(lo.shape[0]):
        upper = x < hi[k] if open_upper else x <= hi[k]
        inl = np.all(np.logical_and(lo[k] < x, upper), axis=1)
        dens += np.where(inl, w[k] / np.prod(hi[k] - lo[k]), 0.0)
    with np.errstate(divide="ignore"):
        return np.log(dens)


# ----------------------------------------------------------------------------------------------
# importance weights
# ----------------------------------------------------------------------------------------------
def dm_weights(target_prob, x, proposal_means, sigma):
    """sampling_methods/dm_ais.py:72-76 (and layered_ais.py:90-94): w = pi(x) / ((1/N) sum_n N(x; mu_n, sigma I)),
    linear domain, one sample at a time in the reference."""
    proposal_means = np.asarray(proposal_means, dtype=np.float64)
    N, dims = proposal_means.shape
    covs = np.stack([np.eye(dims) * sigma] * N)
    q = mixture_prob(x, proposal_means, covs, np.full(N, 1.0 / N))
    with np.errstate(divide="ignore", invalid="ignore"):
        return target_prob(x) / q


def normalize_weights(w):
    """dm_ais.py:86: weights / weights.sum()."""
    w = np.asarray(w, dtype=np.float64)
    return w / w.sum()


def approx_ness(w):
    """sampling_methods/base.py:72-75: (1 / sum w~^2) / n."""
    wn = normalize_weights(w)
    return 1.0 / np.sum(wn * wn) / len(wn)


def mpmc_weights_rho(target_log_prob, x, means, covs, alphas):
    """sampling_methods/m_pmc.py:93-108: w_i = exp(log pi - log q), rho_di = exp(log alpha_d + log q_d - log q);
    weights normalised over the batch."""
    means = np.asarray(means, dtype=np.float64)
    N = means.shape[0]
    lq = mixture_log_prob(x, means, covs, alphas)
    w = np.exp(target_log_prob(x) - lq)
    rho = np.empty((N, x.shape[0]))
    with np.errstate(divide="ignore"):
        for d in range(N):
            rho[d] = np.exp(np.log(alphas[d]) + mvn_log_prob(x, means[d], covs[d]) - lq)
    return w / w.sum(), rho, lq


def mpmc_adapt(x, w, rho, ndims):
    """m_pmc.py:53-78. Note the covariance (:64-68): the UNWEIGHTED scatter (X-mu)^T (X-mu) times
    sum_i w_i rho_di / alpha_d (= 1), with the 10*I failsafe (:71-72)."""
    N = rho.shape[0]
    alpha = np.empty(N)
    mus = np.empty((N, ndims))
    covs = np.empty((N, ndims, ndims))
    for d in range(N):
        alpha[d] = np.sum(w * rho[d])
        with np.errstate(divide="ignore", invalid="ignore"):
            mu = np.sum(w.reshape(-1, 1) * rho[d].reshape(-1, 1) * x, axis=0) / alpha[d]
            scatter = (x - mu).T @ (x - mu)
            cov = scatter * np.sum(w * rho[d]) / alpha[d]
        if np.any(np.isnan(cov)) or np.any(cov > 10):
            cov = np.diag(np.ones(ndims) * 10)
        mus[d], covs[d] = mu, cov
    return alpha / alpha.sum(), mus, covs


def lais_mh(target_prob, means, delta, u, lo, hi):
    """sampling_methods/layered_ais.py:54-74 with the random draws made explicit: delta[l, n] is the MH
    proposal step of chain n at step l (reference: self.mhproposal.sample()[0]), u[l, n] the uniform
    (np.random.rand()). Accept when pi(x_hat) / pi(x) > u (linear-domain ratio)."""
    x = np.array(means, dtype=np.float64)
    L, N = u.shape
    for n in range(N):
        for l in range(L):
            old = target_prob(x[n])[0]
            x_hat = np.clip(x[n] + delta[l, n], lo, hi)
            new = target_prob(x_hat)[0]
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = new / old
            if ratio > u[l, n]:
                x[n] = x_hat
    return x


# ----------------------------------------------------------------------------------------------
# KL divergence
# ----------------------------------------------------------------------------------------------
def kld_terms_from_log_probs(lp, lq):
    """metrics/divergences.py:136-153, on copies (the reference normalises its inputs in place)."""
    lp = np.array(lp, dtype=np.float64)
    lq = np.array(lq, dtype=np.float64)
    inl = np.logical_and(lp > -np.inf, lq > -np.inf)
    lp[inl] -= np.logaddexp.reduce(lp[inl])
    lq[inl] -= np.logaddexp.reduce(lq[inl])
    res = np.exp(lp[inl]) * (lp[inl] - lq[inl])
    res[np.isnan(lq[inl])] = np.inf
    return res


def kld_from_log_probs(lp, lq):
    """divergences.py:105: value = np.sum(compute_from_log_probs(...)) (0.0 for an empty inlier set)."""
    return float(np.sum(kld_terms_from_log_probs(lp, lq)))


def kld(p_log_prob, q_log_prob, samples):
    """divergences.py:101-108."""
    return kld_from_log_probs(p_log_prob(samples), q_log_prob(samples))


def mvn_kld_closed_form(m0, S0, m1, S1):
    """Closed-form KL(N0 || N1) used as the known answer by the reference's
    tests/metrics/test_metric_kldivergence.py:40-47."""
    m0, m1 = np.asarray(m0, dtype=np.float64), np.asarray(m1, dtype=np.float64)
    S0, S1 = np.asarray(S0, dtype=np.float64), np.asarray(S1, dtype=np.float64)
    k = m0.shape[0]
    iS1 = np.linalg.inv(S1)
    diff = m1 - m0
    return 0.5 * (np.trace(iS1 @ S0) + diff @ iS1 @ diff - k + np.log(np.linalg.det(S1) / np.linalg.det(S0)))


# ----------------------------------------------------------------------------------------------
# JSD and EVMSE
# ----------------------------------------------------------------------------------------------
LOG_TINY = -745.13  # exp() of anything below is 0.0 in float64: the log-domain image of the reference's "prob > 0"


def jsd_from_probs(p, q):
    """metrics/divergences.py:165-195 on linear-domain probabilities: inliers = both > 0 and not NaN (:170-173),
    normalise each side over the inliers (:180-181), m = (p + q) / 2 renormalised (:183-186),
    0.5 p log(p/m) + 0.5 q log(q/m) (:188-189), NaN terms dropped (:192), divided by log 2 (:194-195)."""
    p = np.array(p, dtype=np.float64)
    q = np.array(q, dtype=np.float64)
    with np.errstate(all="ignore"):
        inl = (p > 0) & (q > 0) & ~np.isnan(q) & ~np.isnan(p)
        pn = p[inl] / np.sum(p[inl])
        qn = q[inl] / np.sum(q[inl])
        m = 0.5 * (pn + qn)
        m = m / np.sum(m)
        res = 0.5 * pn * np.log(pn / m) + 0.5 * qn * np.log(qn / m)
        res = res[~np.isnan(res)]
    return float(res.sum() / np.log(2))


def jsd_from_log_probs(lp, lq):
    """The same quantity from log densities (what the device path computes; the reference calls p.prob / q.prob, whose
    underflow to 0.0 is the inlier threshold LOG_TINY here): a = lp - LSE_I(lp), b = lq - LSE_I(lq),
    log m = logaddexp(a, b) - log 2, term = 0.5 e^a (a - log m) + 0.5 e^b (b - log m)."""
    lp = np.asarray(lp, dtype=np.float64)
    lq = np.asarray(lq, dtype=np.float64)
    with np.errstate(all="ignore"):
        inl = (lp > LOG_TINY) & (lq > LOG_TINY)
        if not inl.any():
            return 0.0
        a = lp[inl] - np.logaddexp.reduce(lp[inl])
        b = lq[inl] - np.logaddexp.reduce(lq[inl])
        lm = np.logaddexp(a, b) - np.log(2)
        res = 0.5 * np.exp(a) * (a - lm) + 0.5 * np.exp(b) * (b - lm)
        res = res[~np.isnan(res)]
    return float(res.sum() / np.log(2))


def expected_values(x, p, q):
    """metrics/statistics.py:39-45: self-normalised expectations of the evaluation points under p and q."""
    x = np.asarray(x, dtype=np.float64)
    p = np.array(p, dtype=np.float64)
    q = np.array(q, dtype=np.float64)
    with np.errstate(all="ignore"):
        p = p / np.sum(p)
        q = q / np.sum(q)
    return np.sum(x * p.reshape(-1, 1), axis=0), np.sum(x * q.reshape(-1, 1), axis=0)


def evmse(x, p, q):
    """metrics/statistics.py:46-48: squared distance between the two expected values."""
    ev_p, ev_q = expected_values(x, p, q)
    d = ev_p - ev_q
    return float(d @ d)
