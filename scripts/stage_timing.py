"""Per-stage device timing of one KDE-KLD step (debug aid for bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device, parallel
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods.base import CDeviceKDE

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=scale)
dev = torch.device("cuda")
target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
cd, xd = torch.from_numpy(centres_h).to(dev), torch.from_numpy(x_h).to(dev)
n_c = len(centres_h)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e = [ev()]
    kde = CDeviceKDE(cd, None, n_c, 0.02, 8, sup); e.append(ev())
    lq = kde.log_prob(xd); e.append(ev())
    lp = target.log_prob(xd); e.append(ev())
    rec = _device.kld_partials(lp, lq); e.append(ev())
    kld = parallel.global_kld(rec); e.append(ev())
    torch.cuda.synchronize(); w1 = time.perf_counter()
    names = ["pack", "kde_logpdf", "target_logpdf", "kld_partials", "finalize"]
    print(it, "wall %.2f ms |" % ((w1 - w0) * 1e3), " ".join("%s %.3f" % (n, e[i].elapsed_time(e[i + 1])) for i, n in enumerate(names)), "kld", kld)
