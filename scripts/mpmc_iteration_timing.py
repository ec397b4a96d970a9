"""Stage timings of M-PMC iterations at the BASELINE configs[2] shape (8-D 16-mode target, N = 256 proposals, K = 1e6
draws per iteration) with the parameter update on the device vs on the host. Usage: mpmc_iteration_timing.py [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods import CMixturePMC

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
D, N, K = 8, 256, 1_000_000
mu, sig, w, sup = bench.random_gmm(30, D, 16)
target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
for dev_update in (True, False):
    np.random.seed(0); _device.seed(0)
    s = CMixturePMC({"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000, "sigma": 0.05,
                     "device_outputs": True, "device_update": dev_update})
    s.importance_sample(target, K)            # iteration 1 (initial isotropic proposals), also warms everything up
    torch.cuda.synchronize()
    walls = []
    for it in range(iters):
        t0 = time.perf_counter()
        s.importance_sample(target, s._n() + K)
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - t0) * 1e3)
    print("device_update=%s: iteration walls (ms): %s  median %.3f" % (dev_update, ["%.2f" % v for v in walls], float(np.median(walls))))
    # stages of one more iteration, each synchronised
    def tm(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, (time.perf_counter() - t0) * 1e3
    xs, t_s = tm(lambda: s.sample_batch())
    (logw, logq, part), t_w = tm(lambda: s.weigh(xs, target))
    if dev_update:
        _, t_a = tm(lambda: s.adapt_device(xs, logw, logq, part, K))
    else:
        ph = part.cpu().numpy()
        _, t_a = tm(lambda: s.adapt(xs, logw, logq, ph))
    _, t_app = tm(lambda: s._append(xs, logw))
    print("   stages (ms): sample %.3f  weigh %.3f  adapt %.3f  append %.3f   (n accumulated %d)" % (t_s, t_w, t_a, t_app, s._n()))
