"""Time the KDE pairwise kernel alone (CUDA events) for the library selected by AISB_LIB. Usage: kde_kernel_time.py [scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.3
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=scale)
cd, xd = torch.from_numpy(centres_h).cuda(), torch.from_numpy(x_h).cuda()
mix = _device.DeviceMixture.kde(cd, None, len(centres_h), 0.02, spatial_order=os.environ.get("AISB_NO_SORT") != "1")
out = torch.empty(len(x_h), device="cuda")
for _ in range(3):
    _device.mixture_logpdf(xd, mix, out)
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _device.mixture_logpdf(xd, mix, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
pairs = float(len(centres_h)) * len(x_h)
print("sorted=%s" % mix.spatially_ordered, end=" ")
print("%-40s best %.3f ms  median %.3f ms  -> %.4e pairs/s  (checksum %.6f)" % (os.path.basename(os.environ.get("AISB_LIB", "default")), min(ts), float(np.median(ts)), pairs / (min(ts) * 1e-3), float(out.double().mean())))
