"""Time a few pairwise-kernel instantiations for the library selected by AISB_LIB (kernel tuning experiments)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ais_benchmarks_b200 import _device, _lib

def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

rs = np.random.RandomState(0)
out = []
for D in (32, 16):
    n = 2_000_000
    x = torch.from_numpy(rs.uniform(-1, 1, size=(n, D)).astype(np.float32)).cuda()
    res = torch.empty(n, device="cuda")
    for kind, K in (("iso", 1024), ("diag", 64), ("diag", 1024), ("full", 64)):
        mu = rs.uniform(-1, 1, size=(K, D))
        if kind == "iso":
            mix = _device.DeviceMixture.kde(torch.from_numpy(mu).cuda().float().contiguous(), None, K, 0.05)
        elif kind == "diag":
            mix = _device.DeviceMixture.from_params(mu, rs.uniform(0.04, 0.05, size=(K, D)), None, _lib.COV_DIAG)
        else:
            a = rs.normal(size=(K, D, D)) * 0.1
            mix = _device.DeviceMixture.from_params(mu, a @ np.transpose(a, (0, 2, 1)) + 0.05 * np.eye(D), None, _lib.COV_FULL)
        ms = t(lambda: _device.mixture_logpdf(x, mix, res))
        out.append("%s D=%d K=%d: %.3f ms (%.3e pairs/s) chk %.6f" % (kind, D, K, ms, n * K / ms * 1e3, float(res.double().mean())))
print(os.path.basename(os.environ.get("AISB_LIB", "default")))
print("\n".join("  " + o for o in out))
