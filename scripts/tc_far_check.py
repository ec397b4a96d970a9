"""Debug aid: tensor-core log q on data far from the origin vs the CUDA-core kernel and the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ais_benchmarks_b200 import _device
from oracle import ais_oracle as O
for shift in (0.0, 2.0, 8.0):
    rs = np.random.RandomState(11)
    D, N, per = 32, 200, 9
    pm = shift + rs.uniform(-1, 1, size=(N, D))
    pm_d = torch.from_numpy(pm).cuda().float().contiguous()
    mix = _device.DeviceMixture.kde(pm_d, None, N, 0.05)
    _device.seed(5)
    xs = _device.sample_iso(pm_d, per, 0.05)
    hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
    op = _device.TensorCoreOperand(mix)
    ref64 = O.iso_mixture_log_prob(xs.double().cpu().numpy(), pm_d.double().cpu().numpy(), 0.05)
    tc = _device.tc_logq(xs, op, hint).double().cpu().numpy()
    cc = _device.mixture_logpdf(xs, mix).double().cpu().numpy()
    den = np.maximum(1, np.abs(ref64))
    print("shift", shift, "screen_error", op.screen_error, "tc", np.max(np.abs(tc - ref64) / den), "cuda-core",
          np.max(np.abs(cc - ref64) / den), "tc vs cc", np.max(np.abs(tc - cc)), "nan", np.isnan(tc).sum())
