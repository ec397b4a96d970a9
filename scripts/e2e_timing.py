"""Where does the host-buffer (e2e) KDE step spend its time? (debug aid for bench.py)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device, parallel
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods.base import CDeviceKDE

mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=1.0)
dev = torch.device("cuda")
target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
cp, xp = torch.from_numpy(centres_h).pin_memory(), torch.from_numpy(x_h).pin_memory()
n_c = len(centres_h)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(4):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e = [ev()]
    cd = cp.to(dev, non_blocking=True); xd = xp.to(dev, non_blocking=True); e.append(ev()); h1 = time.perf_counter()
    kde = CDeviceKDE(cd, None, n_c, 0.02, 8, sup); e.append(ev()); h2 = time.perf_counter()
    lq = kde.log_prob(xd); e.append(ev()); h3 = time.perf_counter()
    lp = target.log_prob(xd); e.append(ev())
    rec = _device.kld_partials(lp, lq); e.append(ev()); h4 = time.perf_counter()
    kld = parallel.global_kld(rec); e.append(ev())
    torch.cuda.synchronize(); w1 = time.perf_counter()
    names = ["h2d", "pack", "kde_logpdf", "target_logpdf", "kld_partials", "finalize"]
    print(it, "wall %.2f ms |" % ((w1 - w0) * 1e3), " ".join("%s %.3f" % (n, e[i].elapsed_time(e[i + 1])) for i, n in enumerate(names)),
          "| host: h2d-call %.2f pack-call %.2f logpdf-call %.2f rest %.2f" % ((h1 - w0) * 1e3, (h2 - h1) * 1e3, (h3 - h2) * 1e3, (h4 - h3) * 1e3))
