"""Time of aisb_mixture_logpdf_fast_f32 on a rank's share of the configs[3] evaluation points as a function of the number
of interleaved component splits (AISB_SCREEN_NSPLIT; 0 = the library's own choice).
The share is rank 0's of a `1 / fraction`-rank job (parallel.spatial_shard: block-cyclic along the Z-order curve).
Usage: kde_split_sweep.py [point fraction] [splits,comma,separated] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ais_benchmarks_b200 import _device, _lib, parallel
from ais_benchmarks_b200._device import _ptr, stream_ptr, check

frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
splits = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,8,16,24,32,48,64").split(",")]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.load()
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=1.0)
cd, xd = torch.from_numpy(centres_h).cuda(), torch.from_numpy(x_h).cuda()
xs = parallel.spatial_shard(xd, 0, max(1, int(round(1.0 / frac))))[0].contiguous()   # already in Z-order
n, K = xs.shape[0], len(centres_h)
mix = _device.DeviceMixture.kde(cd, None, K, 0.02)
ws = _device.workspace_for(n, K, 8)
sc = _device._kde_fast_scratch(n)
out = torch.empty(n, device="cuda")
stats = torch.zeros(3, dtype=torch.int64, device="cuda")
ref = None
for s in splits:
    if s:
        os.environ["AISB_SCREEN_NSPLIT"] = str(s)
    else:
        os.environ.pop("AISB_SCREEN_NSPLIT", None)
    f = lambda: check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(), _ptr(sc), sc.numel(), None, stream_ptr()))
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ref is None:
        ref = out.clone()
    dev = float(((out - ref).abs() / ref.abs().clamp(min=1)).max())
    stats.zero_()
    check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(), _ptr(sc), sc.numel(), _ptr(stats), stream_ptr()))
    torch.cuda.synchronize()
    print("n=%d splits=%2d  %8.3f ms  (%.4e pairs/s; %.3f ms per 1e11 pairs)  max dev vs first %.2e  listed %d"
          % (n, s, min(ts), float(n) * K / (min(ts) * 1e-3), min(ts) / (float(n) * K / 1e11), dev, int(stats[2])), flush=True)
