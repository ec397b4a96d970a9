"""Where the host-side milliseconds of one e2e KDE-KLD step go (numpy float64 in, KLD out; bench.py's `e2e` leg at
N = 1): wall-clock marks at the entry of the pairwise evaluation, plus a cProfile of one step."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.metrics.divergences import CKLDivergence
from ais_benchmarks_b200.sampling_methods.base import CMixtureSamplingMethod
from ais_benchmarks_b200.distributions import CGaussianMixtureModel

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
c = bench.KDE_CFG
mu, sig, w, sup, centres_h, x_h = bench.kde_workload(scale=scale)
D = c["D"]
target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
centres_np, x_np = centres_h.astype(np.float64), x_h.astype(np.float64)


class _KdeSampler(CMixtureSamplingMethod):
    def importance_sample(self, target_d, n_samples, timeout):
        raise NotImplementedError


marks = {}
orig = _device.mixture_logpdf


def marked(x, mix, *a, **k):
    if mix.K > 1000:
        torch.cuda.synchronize(); marks["enter"] = time.perf_counter()
        out = orig(x, mix, *a, **k)
        torch.cuda.synchronize(); marks["leave"] = time.perf_counter()
        return out
    return orig(x, mix, *a, **k)


def step():
    np.random.seed(1234)
    q = _KdeSampler({"space_min": sup[0], "space_max": sup[1], "dims": D, "n_samples_kde": len(centres_np)})
    q.bw = c["bw"]
    t0 = time.perf_counter()
    q.samples = centres_np
    torch.cuda.synchronize(); marks["samples_set"] = time.perf_counter() - t0
    return CKLDivergence().compute_from_samples(target, q, x_np)


for _ in range(2):
    step()
torch.cuda.synchronize()
_device.mixture_logpdf = marked
import ais_benchmarks_b200.sampling_methods.base as B
for _ in range(2):
    t0 = time.perf_counter(); v = step(); t1 = time.perf_counter()
    print("step %.2f ms: samples setter %.2f | up to the KDE evaluation %.2f | KDE evaluation (sort, kernels, scatter) %.2f | after it %.2f"
          % ((t1 - t0) * 1e3, marks["samples_set"] * 1e3, (marks["enter"] - t0) * 1e3,
             (marks["leave"] - marks["enter"]) * 1e3, (t1 - marks["leave"]) * 1e3))
_device.mixture_logpdf = orig
# the uploads alone: allocation vs the staged copy, per thread count
lib = _device._lib.load()
for th in (2, 4, 8):
    for name, arr in (("centres", centres_np), ("points", x_np)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = torch.empty(arr.shape, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize(); t1 = time.perf_counter()
        _device.check(lib.aisb_upload_as_f32(arr.ctypes.data, 1, arr.size, _device._ptr(out), th, _device.stream_ptr()))
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("upload %-8s %d threads: torch.empty %.2f ms, aisb_upload_as_f32 %.2f ms (%.1f GB/s of float64)"
              % (name, th, (t1 - t0) * 1e3, (t2 - t1) * 1e3, arr.nbytes / (t2 - t1) / 1e9))
        del out
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18); print(s.getvalue()[:3500])
