import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS, CLayeredAIS
for (D, modes, N, K, name) in [(2, 4, 64, 1563, "C2 DM-AIS 2-D, 64 x 1563"), (1, 2, 5, 10, "C1 DM-AIS 1-D, 5 x 10"), (8, 16, 256, 40, "8-D, 256 x 40")]:
    mu, sig, w, sup = bench.random_gmm(20, D, modes)
    target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
    for cls, extra in ((CDeterministicMixtureAIS, {}), (CLayeredAIS, {"L": 10, "mh_sigma": 0.5})):
        for mode in (False, True):
            np.random.seed(0); _device.seed(0)
            s = cls(dict({"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000, "sigma": 0.05,
                          "device_outputs": True, "cuda_graph": mode}, **extra))
            s.importance_sample(target, 2 * N * K)       # warm-up / capture
            s.reset()
            iters = 50
            torch.cuda.synchronize(); t0 = time.perf_counter()
            s.importance_sample(target, iters * N * K)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print("%-28s %-26s cuda_graph=%-5s %.3f ms / iteration (%.3e samples/s), NESS %.4f" % (name, cls.__name__, mode, dt / iters * 1e3, iters * N * K / dt, s.get_approx_NESS()))
