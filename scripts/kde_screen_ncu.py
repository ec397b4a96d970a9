"""Minimal driver for ncu captures of the fast (recentred expanded form) KDE kernel: setup + 3 launches on the configs[3] workload at
`scale` (both sides in Z-order). Usage: kde_screen_ncu.py [scale]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ais_benchmarks_b200 import _device, _lib
from ais_benchmarks_b200._device import _ptr, stream_ptr, check

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
lib = _lib.load()
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=scale)
cd, xd = torch.from_numpy(centres_h).cuda(), torch.from_numpy(x_h).cuda()
n, K = len(x_h), len(centres_h)
mix = _device.DeviceMixture.kde(cd, None, K, 0.02, spatial_order=True)
if not mix.spatially_ordered:       # below MORTON_MIN_ROWS the library does not sort: do it here
    cd = cd[_device.morton_order(cd)].contiguous()
    mix = _device.DeviceMixture.kde(cd, None, K, 0.02, spatial_order=False)
xs = xd[_device.morton_order(xd)].contiguous()
ws = _device.workspace_for(n, K, 8)
out = torch.empty(n, device="cuda")
sc = _device._kde_fast_scratch(n)
for _ in range(3):
    check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(), _ptr(sc), sc.numel(), None, stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.double().mean()))
