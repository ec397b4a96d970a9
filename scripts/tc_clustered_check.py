"""Tensor-core proposal density on CLUSTERED (post-adaptation) DM-PMC proposals: accuracy of the tcgen05 path and of
the CUDA-core kernel against the float64 oracle, and the time of both. Usage: tc_clustered_check.py [D] [modes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from oracle import ais_oracle as O
from ais_benchmarks_b200 import _device
from ais_benchmarks_b200.distributions import CGaussianMixtureModel
from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
modes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N, K = 1024, 40
mu, sig, w, sup = bench.random_gmm(10 + D, D, modes)
target = CGaussianMixtureModel({"means": mu, "sigmas": sig, "weights": w, "support": sup})
np.random.seed(5); _device.seed(5)
s = CDeterministicMixtureAIS({"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000,
                              "sigma": 0.05, "cuda_graph": False, "device_outputs": True})
s.importance_sample(target, 5 * N * K)
pm = s.proposal_means().clone()
print("D=%d: %d distinct proposal means of %d after 5 adaptations, NESS %.2e" % (D, torch.unique(pm, dim=0).shape[0], N, s.get_NESS()))
qmix = _device.DeviceMixture.kde(pm, None, N, 0.05)
op = _device.tc_operand(qmix)
per = 300
xs = _device.sample_iso(pm, per, 0.05)
hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
got = _device.tc_logq(xs, op, hint)
ref32 = _device.mixture_logpdf(xs, qmix)
rows = np.random.RandomState(D).choice(xs.shape[0], size=2048, replace=False)
ref64 = O.iso_mixture_log_prob(xs[rows].double().cpu().numpy(), pm.double().cpu().numpy(), 0.05)
rel = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(1, np.abs(b))))
print("tc   vs float64 oracle (2048 rows): %.3e" % rel(got[rows].double().cpu().numpy(), ref64))
print("cuda vs float64 oracle (2048 rows): %.3e" % rel(ref32[rows].double().cpu().numpy(), ref64))
d = ((got.double() - ref32.double()).abs() / ref32.double().abs().clamp(min=1.0))
print("tc vs cuda-core on all %d rows: max %.3e, rows above 2e-6: %d, above 5e-6: %d" % (xs.shape[0], float(d.max()), int((d > 2e-6).sum()), int((d > 5e-6).sum())))
i = int(d.argmax())
print("worst row %d: tc %.7f cuda %.7f oracle %.7f" % (i, float(got[i]), float(ref32[i]), float(O.iso_mixture_log_prob(xs[i:i+1].double().cpu().numpy(), pm.double().cpu().numpy(), 0.05)[0])))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
out = torch.empty_like(got)
pm0 = torch.from_numpy(np.random.RandomState(1).uniform(sup[0], sup[1], size=(N, D))).cuda().float().contiguous()
q0 = _device.DeviceMixture.kde(pm0, None, N, 0.05); x0 = _device.sample_iso(pm0, per, 0.05)
print("time for %d samples x %d proposals: clustered tc %.3f ms, cuda-core %.3f ms | un-adapted tc %.3f ms, cuda-core %.3f ms" % (
    xs.shape[0], N, timed(lambda: _device.tc_logq(xs, op, hint, out)), timed(lambda: _device.mixture_logpdf(xs, qmix, out)),
    timed(lambda: _device.tc_logq(x0, _device.tc_operand(q0), hint, out)), timed(lambda: _device.mixture_logpdf(x0, q0, out))))
