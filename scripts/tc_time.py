"""Time tc_logq_kernel alone at the DM config-5 shape (AISB_TC_DEBUG selects the timing experiments). Usage: tc_time.py [scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
mu2, sig2, w2, sup2, pmeans, per = bench.dm_workload(scale=scale)
N = len(pmeans)
pm_d = torch.from_numpy(pmeans).cuda().float().contiguous()
mix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
_device.seed(3)
xs = _device.sample_iso(pm_d, per, 0.05)
hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
op = _device.TensorCoreOperand(mix)
out = torch.empty(xs.shape[0], device="cuda")
_device.tc_logq(xs, op, hint, out); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _device.tc_logq(xs, op, hint, out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("AISB_TC_DEBUG=%s  %d samples x %d proposals: %.3f ms -> %.3e pairs/s (checksum %.6f)" % (
    os.environ.get("AISB_TC_DEBUG", "0"), xs.shape[0], N, min(ts), xs.shape[0] * N / (min(ts) * 1e-3), float(out.double().mean())))
