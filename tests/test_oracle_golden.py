"""Pins the oracle (oracle/ais_oracle.py) to the reference: every committed fixture produced by the
unmodified reference (tests/golden/make_golden.py) and the known answers of the reference's own tests.
CPU only."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from conftest import load_golden
from oracle import ais_oracle as O

TOL = 1e-11


def close(a, b, tol=TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(a[~fin], b[~fin], equal_nan=True)
    assert np.allclose(a[fin], b[fin], rtol=tol, atol=tol), np.max(np.abs(a[fin] - b[fin]))


def test_mvn_matches_reference_and_scipy():
    g = load_golden("mvn")
    for i in range(int(g["n"])):
        mean, cov, x = g["mean%d" % i], g["cov%d" % i], g["x%d" % i]
        close(O.mvn_log_prob(x, mean, cov), g["logp%d" % i])
        close(O.mvn_prob(x, mean, cov), g["prob%d" % i])
        close(O.mvn_default_support(mean, cov), g["support%d" % i])
    # reference tests/distributions/test_multivariate_normal_unittest.py:17-48 known answers
    for i in (0, 1):
        mean, cov, x = g["mean%d" % i], g["cov%d" % i], g["x%d" % i]
        close(O.mvn_prob(x, mean, cov), multivariate_normal.pdf(x, mean=mean, cov=cov))
        close(O.mvn_prob(x[0], mean, cov), [multivariate_normal.pdf(x[0], mean=mean, cov=cov)])


def test_gmm_matches_reference():
    g = load_golden("gmm")
    for i in range(int(g["n"])):
        x = g["x%d" % i]
        close(O.gmm_log_prob(x, g["means%d" % i], g["sigmas%d" % i], g["weights%d" % i]), g["logp%d" % i])
        close(O.gmm_prob(x, g["means%d" % i], g["sigmas%d" % i], g["weights%d" % i]), g["prob%d" % i], 1e-10)


def test_mixture_weight_sanitising_and_full_cov():
    g = load_golden("mixture")
    close(O.sanitize_weights(g["weights"]), g["norm_weights"])
    close(O.mixture_log_prob(g["x"], g["means"], g["covs"], g["weights"]), g["logp"])
    close(O.mixture_prob(g["x"], g["means"], g["covs"], g["weights"]), g["prob"])
    close(O.mixture_log_prob(g["x"][0], g["means"], g["covs"], g["weights"]), g["logp_single"])


def test_dm_weights_match_reference_loop():
    g = load_golden("dm_weights")
    for i in range(int(g["n"])):
        x, pm, sigma = g["x%d" % i], g["pmeans%d" % i], float(g["sigma%d" % i])
        tm, ts, tw = g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i]
        w = O.dm_weights(lambda s: O.gmm_prob(s, tm, ts, tw), x, pm, sigma)
        close(w, g["w%d" % i], 1e-10)
        close(O.normalize_weights(w), g["norm_w%d" % i], 1e-10)
        close(O.approx_ness(w), g["ness%d" % i], 1e-10)
        close(O.iso_mixture_log_prob(x, pm, sigma), g["logq%d" % i], 1e-10)
        close(O.gmm_log_prob(x, tm, ts, tw), g["logp%d" % i])


def test_kde_and_kld_match_reference():
    g = load_golden("kde_kld")
    for i in range(int(g["n"])):
        x, centres, bw = g["x%d" % i], g["centres%d" % i], float(g["bw%d" % i])
        lq = O.iso_mixture_log_prob(x, centres, bw)
        lp = O.gmm_log_prob(x, g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i])
        close(lq, g["lq%d" % i], 1e-10)
        close(lp, g["lp%d" % i])
        close(O.kld_terms_from_log_probs(lp, lq), g["terms%d" % i], 1e-9)
        assert abs(O.kld_from_log_probs(lp, lq) - float(g["kld%d" % i])) <= 1e-9 * max(1.0, abs(float(g["kld%d" % i])))
    assert O.kld_from_log_probs(g["edge_lp"], g["edge_lq"]) == pytest.approx(float(g["edge_kld"]), abs=1e-13)
    assert O.kld_from_log_probs(np.full(4, -np.inf), np.zeros(4)) == float(g["empty_kld"]) == 0.0


def test_jsd_and_evmse_match_reference():
    """metrics/divergences.py:156-195 and metrics/statistics.py:29-48 on the KDE fixtures, a Gaussian pair and the
    zero / NaN edge cases; the log-domain JSD (what the device computes) must agree with the linear-domain one."""
    g, k = load_golden("metrics"), load_golden("kde_kld")
    with np.errstate(all="ignore"):
        for i in range(int(g["n"])):
            x, lp, lq = k["x%d" % i], k["lp%d" % i], k["lq%d" % i]
            pp = O.gmm_prob(x, k["tmeans%d" % i], k["tsigmas%d" % i], k["tweights%d" % i])
            qq = np.exp(lq)
            assert O.jsd_from_probs(pp, qq) == pytest.approx(float(g["jsd%d" % i]), rel=1e-9)
            assert O.jsd_from_log_probs(lp, lq) == pytest.approx(float(g["jsd%d" % i]), rel=1e-9)
            assert O.evmse(x, pp, qq) == pytest.approx(float(g["evmse%d" % i]), rel=1e-9)
            ev_p, ev_q = O.expected_values(x, pp, qq)
            close(ev_p, g["ev_p%d" % i], 1e-10)
            close(ev_q, g["ev_q%d" % i], 1e-10)
        xn = g["norm_x"]
        pn, qn = O.mvn_prob(xn, np.array([0.0]), np.eye(1)), O.mvn_prob(xn, np.array([1.0]), np.eye(1))
        assert O.jsd_from_probs(pn, qn) == pytest.approx(float(g["norm_jsd"]), rel=1e-10)
        assert O.jsd_from_log_probs(np.log(pn), np.log(qn)) == pytest.approx(float(g["norm_jsd"]), rel=1e-10)
        assert O.evmse(xn, pn, qn) == pytest.approx(float(g["norm_evmse"]), rel=1e-10)
        assert O.jsd_from_probs(g["edge_p"], g["edge_q"]) == pytest.approx(float(g["edge_jsd"]), rel=1e-12)
        assert O.jsd_from_log_probs(np.log(g["edge_p"]), np.log(g["edge_q"])) == pytest.approx(float(g["edge_jsd"]), rel=1e-12)
        assert O.evmse(g["edge_x"], g["edge_fp"], g["edge_fq"]) == pytest.approx(float(g["edge_evmse"]), rel=1e-12)
    assert O.jsd_from_log_probs(np.full(3, -np.inf), np.zeros(3)) == 0.0


def test_kld_known_answers_of_reference_tests():
    """tests/metrics/test_metric_kldivergence.py:77-135: within 1 % of the closed-form Gaussian KL on 1e5 uniform
    points; and identical (to rounding) to what the reference itself returned on the same seeded points."""
    g = load_golden("kde_kld")
    for j in range(int(g["n_kat"])):
        m0, m1, s0, s1 = g["kat_m0_%d" % j], g["kat_m1_%d" % j], g["kat_s0_%d" % j], g["kat_s1_%d" % j]
        np.random.seed(int(g["kat_seed_%d" % j]))
        sup = O.mvn_default_support(m0, s0)
        x = np.random.uniform(sup[0], sup[1], size=(100000, len(m0)))
        val = O.kld(lambda s: O.mvn_log_prob(s, m0, s0), lambda s: O.mvn_log_prob(s, m1, s1), x)
        assert val == pytest.approx(float(g["kat_kld_%d" % j]), rel=1e-9)
        expected = O.mvn_kld_closed_form(m0, s0, m1, s1)
        assert abs(val - expected) <= 0.01 * expected


def test_mpmc_iteration_matches_reference():
    g = load_golden("mpmc")
    for i in range(int(g["n"])):
        tm, ts, tw = g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i]
        for it in range(2):
            t = "%d_%d" % (i, it)
            x = g["x" + t]
            w, rho, _ = O.mpmc_weights_rho(lambda s: O.gmm_log_prob(s, tm, ts, tw), x, g["means" + t], g["covs" + t],
                                           g["alpha" + t])
            close(w, g["w" + t], 1e-9)
            alpha, mus, covs = O.mpmc_adapt(x, w, rho, x.shape[1])
            close(alpha, g["new_alpha" + t], 1e-9)
            close(mus, g["new_means" + t], 1e-8)
            close(covs, g["new_covs" + t], 1e-8)


def test_lais_chains_match_reference():
    g = load_golden("lais")
    for i in range(int(g["n"])):
        tm, ts, tw = g["tmeans%d" % i], g["tsigmas%d" % i], g["tweights%d" % i]
        after = O.lais_mh(lambda s: O.gmm_prob(s, tm, ts, tw), g["before%d" % i], g["delta%d" % i], g["u%d" % i],
                          g["lo%d" % i], g["hi%d" % i])
        close(after, g["after%d" % i])
        assert int(g["pi_evals%d" % i]) == 2 * g["u%d" % i].size


def test_uniform_and_box_mixture_match_reference():
    g = load_golden("uniform")
    for i in range(int(g["n"])):
        lp = O.uniform_log_prob(g["x%d" % i], g["c%d" % i], g["r%d" % i])
        close(lp, g["logp%d" % i])
        close(np.exp(lp), g["prob%d" % i])
    # known answer of tests/distributions/test_multivariate_uniform_unittest.py: 1 / prod(2 r) inside
    assert np.exp(O.uniform_log_prob(np.array([[0.0, 0.1]]), [0.0, 0.1], [0.5, 0.2]))[0] == pytest.approx(1 / (1.0 * 0.4))
    c, r, w = g["mix_centers"], g["mix_radii"], g["mix_norm_w"]
    lp = O.box_mixture_log_prob(g["mix_x"], c - r, c + r, w, open_upper=False)
    close(lp, g["mix_logp"], 1e-10)
