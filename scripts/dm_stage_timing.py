"""Per-kernel device time of one DM-AIS weighting step at BASELINE configs[4] size (debug aid for bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device, parallel
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
mu2, sig2, w2, sup2, pmeans, per = bench.dm_workload(scale=scale)
N = len(pmeans)
dev = torch.device("cuda")
target = CGaussianMixtureModel({"means": mu2, "sigmas": sig2, "weights": w2, "support": sup2})
tmix = target.device_mixture()
pm_d = torch.from_numpy(pmeans).to(dev, torch.float32).contiguous()
qmix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
_device.seed(1)
xs = _device.sample_iso(pm_d, per, 0.05)
own = (torch.arange(xs.shape[0], device=dev) // per).to(torch.int32)
op = _device.TensorCoreOperand(qmix)
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
lp = torch.empty(xs.shape[0], device=dev); lq = torch.empty_like(lp)
for it in range(3):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e = [ev()]
    _device.mixture_logpdf(xs, tmix, lp); e.append(ev())
    _device.tc_logq(xs, op, own, lq); e.append(ev())
    torch.cuda.synchronize(); w1 = time.perf_counter()
    e2 = [ev()]
    logw, _, _, part = _device.dm_weights(xs, tmix, None, qmix, hint=own); e2.append(ev())
    g = part.cpu().numpy(); e2.append(ev())
    wn = _device.normalize_weights(logw, g); e2.append(ev())
    torch.cuda.synchronize(); w2_ = time.perf_counter()
    print(it, "target_logpdf %.3f tc_logq(+err sync) %.3f | dm_weights_hinted %.3f partials.cpu %.3f normalize %.3f | walls %.2f %.2f ms" % (
        e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e2[0].elapsed_time(e2[1]), e2[1].elapsed_time(e2[2]), e2[2].elapsed_time(e2[3]), (w1-w0)*1e3, (w2_-w1)*1e3))
