"""Time the two Z-order implementations on 1e6 x 8 points (key kernel + 32-bit sort vs the torch element-wise chain)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ais_benchmarks_b200 import _device
x = torch.rand(1_000_000, 8, device="cuda")
for name, fn in (("kernel + int32 sort", _device.morton_order), ("torch chain + int64 argsort", _device.morton_order_torch)):
    for _ in range(3): fn(x)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("%-30s %.3f ms" % (name, min(ts)))
