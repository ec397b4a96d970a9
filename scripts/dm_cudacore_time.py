import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
mu2, sig2, w2, sup2, pmeans, per = bench.dm_workload(scale=0.5)
N = len(pmeans)
target = CGaussianMixtureModel({"means": mu2, "sigmas": sig2, "weights": w2, "support": sup2})
tmix = target.device_mixture()
pm_d = torch.from_numpy(pmeans).cuda().float().contiguous()
qmix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
_device.seed(1)
xs = _device.sample_iso(pm_d, per, 0.05)
for _ in range(2): _device.dm_weights(xs, tmix, None, qmix)
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); logw, _, _, part = _device.dm_weights(xs, tmix, None, qmix); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.path.basename(os.environ.get("AISB_LIB", "default")), "fused CUDA-core dm_weights, %d samples x %d: %.3f ms, checksum %.6f" % (xs.shape[0], N, min(ts), float(logw.double().mean())))
