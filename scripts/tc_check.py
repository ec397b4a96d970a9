"""Accuracy and speed of the tensor-core cross-term path vs the FP32 path and the float64 oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ais_benchmarks_b200 import _device
from oracle import ais_oracle as O

def check(D, K, n, var, seed, spread=1.0, big_n=0):
    rs = np.random.RandomState(seed)
    mu = rs.uniform(-spread, spread, size=(K, D))
    comp = rs.randint(0, K, size=n)
    x = (mu[comp] + rs.standard_normal((n, D)) * np.sqrt(var)).astype(np.float32)
    x[: n // 4] = rs.uniform(-spread, spread, size=(n // 4, D)).astype(np.float32)
    md = torch.from_numpy(mu).cuda().float().contiguous()
    mix = _device.DeviceMixture.kde(md, None, K, var)
    xd = torch.from_numpy(x).cuda()
    ref32 = _device.mixture_logpdf(xd, mix).double().cpu().numpy()
    op = _device.TensorCoreOperand(mix)
    hint = torch.from_numpy(comp.astype(np.int32)).cuda()
    got = _device.tc_logq(xd, op, hint).double().cpu().numpy()
    sub = rs.choice(n, size=min(n, 512), replace=False)
    ref64 = O.iso_mixture_log_prob(x[sub].astype(np.float64), md.double().cpu().numpy(), var)
    e_tc = np.abs(got[sub] - ref64); e_32 = np.abs(ref32[sub] - ref64)
    rel_tc = e_tc / np.maximum(1, np.abs(ref64)); rel_32 = e_32 / np.maximum(1, np.abs(ref64))
    print("D=%d K=%d n=%d var=%g: TC abs err max %.2e mean %.2e rel(max(1,|ref|)) max %.2e | FP32 path abs max %.2e rel max %.2e | |ref| range [%.1f, %.1f]"
          % (D, K, n, var, e_tc.max(), e_tc.mean(), rel_tc.max(), e_32.max(), rel_32.max(), np.abs(ref64).min(), np.abs(ref64).max()), flush=True)
    if big_n:
        pick = torch.randint(0, n, (big_n,), device="cuda")
        pick, _ = torch.sort(pick)
        xb = xd[pick].contiguous(); hb = hint[pick].contiguous()
        out = torch.empty(big_n, device="cuda")
        for name, fn in (("tc  ", lambda: _device.tc_logq(xb, op, hb, out)), ("fp32", lambda: _device.mixture_logpdf(xb, mix, out))):
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print("   %s: %.3f ms -> %.3e pairs/s" % (name, min(ts), big_n * K / (min(ts) * 1e-3)), flush=True)

stage = sys.argv[1] if len(sys.argv) > 1 else "small"
if stage == "small":
    check(32, 128, 128, 0.05, 0)
    check(32, 1024, 4096, 0.05, 1)
    check(8, 640, 1000, 0.02, 2)
    check(16, 300, 777, 0.05, 3)
else:
    check(32, 1024, 65536, 0.05, 4, big_n=2_000_000)
    check(8, 100_000, 65536, 0.02, 5, big_n=200_000)
    check(16, 4096, 65536, 0.05, 6, big_n=1_000_000)

if stage == "dm":
    mu2, sig2, w2, sup2, pmeans, per = bench.dm_workload(scale=float(sys.argv[2]) if len(sys.argv) > 2 else 0.2)
    N = len(pmeans)
    pm_d = torch.from_numpy(pmeans).cuda().float().contiguous()
    mix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
    _device.seed(3)
    xs = _device.sample_iso(pm_d, per, 0.05)
    hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
    op = _device.TensorCoreOperand(mix)
    out_tc = torch.empty(xs.shape[0], device="cuda"); out_32 = torch.empty_like(out_tc)
    for name, fn in (("tc  ", lambda: _device.tc_logq(xs, op, hint, out_tc)), ("fp32", lambda: _device.mixture_logpdf(xs, mix, out_32))):
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print("DM config-5 shape, %d samples x %d proposals: %s %.3f ms -> %.3e pairs/s" % (xs.shape[0], N, name, min(ts), xs.shape[0] * N / (min(ts) * 1e-3)), flush=True)
    d = (out_tc.double() - out_32.double()).abs()
    rows = np.random.RandomState(0).choice(xs.shape[0], 256, replace=False)
    ref = O.iso_mixture_log_prob(xs[rows].double().cpu().numpy(), pm_d.double().cpu().numpy(), 0.05)
    print("   |tc - fp32| max %.2e ; vs oracle: tc rel %.2e fp32 rel %.2e" % (float(d.max()),
          float(np.max(np.abs(out_tc[rows].double().cpu().numpy() - ref) / np.maximum(1, np.abs(ref)))),
          float(np.max(np.abs(out_32[rows].double().cpu().numpy() - ref) / np.maximum(1, np.abs(ref))))))
