"""BASELINE.json configs[1] and configs[2] through the public sampler / metric API on one GPU
(the adapt loops of DM-AIS, M-PMC and LAIS at the named shapes), with a float64 oracle check of the weights of a
random row subset. Prints one line per stage; `python scripts/config_timings.py > profiles/<name>.txt`.

  C2: DM-AIS on a 2-D 4-mode GMM, 1e5 samples x 64 proposals, KLD over 1e5 evaluation points
  C3: 8-D 16-mode GMM target, M-PMC (N=256, K=1e6 per iteration) and LAIS (N=256, K=3907, L=10), 1e6 samples
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.metrics import CKLDivergence
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS, CMixturePMC
from oracle import ais_oracle as O


def timed(fn, reps=3):
    out, best = None, 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best


def check_weights(name, x, wn, logq_fn, mu, sig, w, rows=2000, seed=0):
    """Normalised weights of a random row subset against the float64 oracle (same samples, same proposal params)."""
    rs = np.random.RandomState(seed)
    x = x.detach().cpu().numpy().astype(np.float64)
    wn = wn.detach().double().cpu().numpy()
    pick = rs.choice(len(x), size=min(rows, len(x)), replace=False)
    ref_lw = O.gmm_log_prob(x[pick], mu, sig, w) - logq_fn(x[pick])
    got_lw = np.log(wn[pick])
    ok = np.isfinite(got_lw) & np.isfinite(ref_lw)
    # normalised weights differ from exp(ref_lw) by one global constant: compare after removing it
    shift = np.median((got_lw - ref_lw)[ok])
    err = np.max(np.abs(got_lw[ok] - shift - ref_lw[ok]) / np.maximum(1.0, np.abs(ref_lw[ok])))
    print("    parity[%s]: %d rows vs float64 oracle, max |dlogw| / max(1,|logw|) = %.2e" % (name, ok.sum(), err))
    assert err < 1e-4, err


def config2():
    D, modes, N = 2, 4, 64
    K = 1563
    mu, sig, w, sup = bench.random_gmm(20, D, modes)
    target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
    params = {"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000, "sigma": 0.05,
              "device_outputs": True}
    print("C2: DM-AIS, 2-D 4-mode GMM, N=%d proposals x K=%d draws = %d samples per iteration" % (N, K, N * K))

    def one_iteration():
        np.random.seed(0)
        _device.seed(0)
        s = CDeterministicMixtureAIS(params)
        pm = s.proposal_means().double().cpu().numpy()
        x, wn = s.importance_sample(target, N * K)
        return s, pm, x, wn
    (s, pm, x, wn), dt = timed(one_iteration)
    print("    importance_sample (1 iteration: sample + weigh + resample + normalise): %.3f ms -> %.3e samples/s, NESS %.4f"
          % (dt * 1e3, N * K / dt, s.get_approx_NESS() if hasattr(s, "get_approx_NESS") else float("nan")))
    check_weights("C2 DM-AIS", x, wn, lambda v: O.iso_mixture_log_prob(v, pm, 0.05), mu, sig, w)

    def ten_iterations():
        np.random.seed(0)
        _device.seed(0)
        s = CDeterministicMixtureAIS(params)
        s.importance_sample(target, 10 * N * K)
        return s
    s, dt = timed(ten_iterations, reps=2)
    print("    importance_sample (10 adapt iterations, %d samples): %.3f ms -> %.3e samples/s" % (s._n(), dt * 1e3, s._n() / dt))
    metric = CKLDivergence()
    np.random.seed(1)
    kld, dt = timed(lambda: metric.compute(p=target, q=s, nsamples=100000))
    print("    CKLDivergence.compute(target, sampler, 1e5 eval points): %.3f ms, KLD %.5f" % (dt * 1e3, kld))


def config3():
    D, modes, N = 8, 16, 256
    mu, sig, w, sup = bench.random_gmm(0, D, modes)
    target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
    K = 1_000_000
    params = {"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000, "sigma": 0.05,
              "device_outputs": True}
    print("C3: M-PMC, 8-D 16-mode GMM, N=%d proposals, K=%d samples per iteration" % (N, K))

    def mpmc(iters):
        np.random.seed(0)
        _device.seed(0)
        s = CMixturePMC(params)
        m0 = (s._means.copy(), np.array([np.diag(c) for c in s._covs]), s._mix_weights.copy())
        x, wn = s.importance_sample(target, iters * K)
        return s, m0, x, wn
    (s, m0, x, wn), dt = timed(lambda: mpmc(1))
    print("    importance_sample (1 iteration: sample + weigh + posterior moments + adapt): %.3f ms -> %.3e samples/s"
          % (dt * 1e3, K / dt))
    check_weights("C3 M-PMC it.1", x[:K], wn[:K], lambda v: O.gmm_log_prob(v, m0[0], m0[1], m0[2]), mu, sig, w)
    (s, _, x, wn), dt = timed(lambda: mpmc(3), reps=2)
    print("    importance_sample (3 iterations): %.3f ms -> %.3e samples/s"
          % (dt * 1e3, s._n() / dt))
    print("      stage timings of one iteration on the adapted proposals:")
    xs, dt = timed(lambda: s.sample_batch())
    print("        sample_batch %.3f ms" % (dt * 1e3))
    (logw, logq, part), dt = timed(lambda: s.weigh(xs, target))
    print("        weigh (log pi, log q over %d adapted proposals [10*I after the Q3 failsafe], log w, record) %.3f ms -> %.3e pairs/s"
          % (N, dt * 1e3, K * N / dt))
    ph = part.cpu().numpy()
    _, dt = timed(lambda: _device.mpmc_moments(xs, s.proposal_mixture(), logq, logw, _device.weight_logsum(ph)))
    print("        mpmc_moments (alpha, mu numerators over %d x %d pairs) %.3f ms" % (K, N, dt * 1e3))

    Kl = 3907
    lp = {"space_min": sup[0], "space_max": sup[1], "dims": D, "K": Kl, "N": N, "J": 1000, "L": 10, "sigma": 0.05,
          "mh_sigma": 0.5, "device_outputs": True}
    print("C3: LAIS, N=%d chains x L=10 MH steps, K=%d draws per proposal = %d samples per iteration" % (N, Kl, N * Kl))

    def lais():
        np.random.seed(0)
        _device.seed(0)
        s = CLayeredAIS(lp)
        x, wn = s.importance_sample(target, N * Kl)
        return s, x, wn
    (s, x, wn), dt = timed(lais)
    print("    importance_sample (1 iteration: MH layer + sample + weigh + normalise): %.3f ms -> %.3e samples/s"
          % (dt * 1e3, s._n() / dt))


if __name__ == "__main__":
    which = sys.argv[1:] or ["2", "3"]
    if "2" in which:
        config2()
    if "3" in which:
        config3()
