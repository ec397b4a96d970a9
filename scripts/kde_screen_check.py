"""Screen-and-refine KDE kernel against the fold-skipping and the plain pairwise kernels on the configs[3] workload:
agreement (max |diff| of the log densities), time per launch (CUDA events) and the refined fraction of the screen.
Usage: kde_screen_check.py [scale] [reps] [point fraction]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device, _lib
from ais_benchmarks_b200._device import _ptr, stream_ptr, check

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=scale)
if len(sys.argv) > 3:                      # evaluate only a fraction of the points (the per-rank share of an N-GPU run)
    x_h = x_h[:int(len(x_h) * float(sys.argv[3]))]
cd, xd = torch.from_numpy(centres_h).cuda(), torch.from_numpy(x_h).cuda()
n, K = len(x_h), len(centres_h)
mix = _device.DeviceMixture.kde(cd, None, K, 0.02)
assert mix.spatially_ordered or K < _device.MORTON_MIN_ROWS
xs = xd[_device.morton_order(xd)].contiguous()
ws = _device.workspace_for(n, K, 8)
screen = _device.kde_screen(mix)


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


o_plain, o_ord, o_scr, o_fast = (torch.empty(n, device="cuda") for _ in range(4))
stats = torch.zeros(3, dtype=torch.int64, device="cuda")
sc = _device._kde_fast_scratch(n)
f_plain = lambda: check(lib.aisb_mixture_logpdf_f32(_ptr(xs), n, mix.ref(), _ptr(o_plain), _ptr(ws), ws.numel(), stream_ptr()))
f_ord = lambda: check(lib.aisb_mixture_logpdf_ordered_f32(_ptr(xs), n, mix.ref(), _ptr(o_ord), _ptr(ws), ws.numel(), stream_ptr()))
f_scr = lambda: check(lib.aisb_mixture_logpdf_screened_f32(_ptr(xs), n, mix.ref(), _ptr(screen), _ptr(o_scr), _ptr(ws), ws.numel(), None, stream_ptr()))
f_cnt = lambda: check(lib.aisb_mixture_logpdf_screened_f32(_ptr(xs), n, mix.ref(), _ptr(screen), _ptr(o_scr), _ptr(ws), ws.numel(), _ptr(stats), stream_ptr()))
f_fast = lambda: check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(o_fast), _ptr(ws), ws.numel(), _ptr(sc), sc.numel(), None, stream_ptr()))
f_fcnt = lambda: check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(o_fast), _ptr(ws), ws.numel(), _ptr(sc), sc.numel(), _ptr(stats), stream_ptr()))
pairs = float(n) * K
for name, f in (("plain", f_plain), ("ordered (fold skip)", f_ord), ("screened", f_scr), ("fast (recentred)", f_fast)):
    t = timed(f)
    print("%-22s %9.3f ms  %.4e pairs/s" % (name, t, pairs / (t * 1e-3)))
stats.zero_(); f_cnt(); torch.cuda.synchronize()
st = stats.cpu().numpy()
print("screen units refined / screened: %d / %d = %.4f" % (st[0], st[1], st[0] / max(1, st[1])))
stats.zero_(); f_fcnt(); torch.cuda.synchronize()
st = stats.cpu().numpy()
print("fast units folded / screened: %d / %d = %.4f ; points re-evaluated exactly: %d of %d" % (st[0], st[1], st[0] / max(1, st[1]), st[2], n))
a, b, c = o_plain.double(), o_ord.double(), o_scr.double()
f = o_fast.double()
rel = (f - a).abs() / torch.clamp(a.abs(), min=1.0)
print("max |fast - plain| / max(1,|plain|)     = %.3e  (points above 3e-6: %d, above 1e-5: %d)" % (float(rel.max()), int((rel > 3e-6).sum()), int((rel > 1e-5).sum())))
a_iso = float(lib.aisb_iso_scale(0.02))
npad = (n + 127) // 128 * 128
xp = torch.cat([xs, xs[-1:].expand(npad - n, -1)]).reshape(-1, 128, xs.shape[1]).double()
cen = xp.mean(1, keepdim=True)
U = (((xp - cen) * a_iso) ** 2).sum(-1).reshape(-1)[:n]
worst = torch.argsort(rel, descending=True)[:8]
for i in worst.tolist():
    print("   i=%7d lp_plain=%10.5f lp_fast=%10.5f rel=%.2e  U_i=%8.2f U_warp_max=%8.2f" % (i, float(a[i]), float(f[i]), float(rel[i]), float(U[i]), float(U[(i // 128) * 128:(i // 128) * 128 + 128].max())))
den = torch.clamp(a.abs(), min=1.0)
print("max |ordered - plain| / max(1,|plain|)  = %.3e" % float(((b - a).abs() / den).max()))
print("max |screened - plain| / max(1,|plain|) = %.3e" % float(((c - a).abs() / den).max()))
print("max |screened - ordered| abs            = %.3e" % float((c - b).abs().max()))
# float64 oracle on a random row subset
from oracle import ais_oracle as O
rs = np.random.RandomState(3)
rows = rs.choice(n, size=min(n, 24), replace=False)
xs_h = xs[rows].cpu().numpy().astype(np.float64)
ref = O.iso_mixture_log_prob(xs_h, centres_h.astype(np.float64), 0.02)
got = o_scr[rows].cpu().numpy().astype(np.float64)
print("screened vs float64 oracle on %d rows: max rel err %.3e" % (len(rows), float(np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))))))
got = o_fast[rows].cpu().numpy().astype(np.float64)
print("fast     vs float64 oracle on %d rows: max rel err %.3e" % (len(rows), float(np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))))))
