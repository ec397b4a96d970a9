"""Wall time of the (target, method) evaluation loop, fused vs the reference's data flow.
usage: python scripts/eval_loop_timing.py [nsamples] [nsamples_eval]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ais_benchmarks_b200.benchmark import CBenchmark
here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ais_benchmarks_b200", "benchmark")
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
b = CBenchmark()
b.load_benchmark(os.path.join(here, "def_benchmark.yaml"))
t = b.targets[1]
b.load_methods(os.path.join(here, "def_methods.yaml"), t.support()[0], t.support()[1], t.dims)
for m in b.methods:
    for fused in (True, False, True, False):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rows = CBenchmark.evaluate_method(t.dims, t, m, max_samples=ns, sampling_eval_samples=ne,
                                          metrics=["KLD", "JSD", "EVMSE"], rseed=1, n_reps=1, batch_size=2, fused=fused)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("%-14s fused=%-5s iterations %4d  %.3f s  (%.2f ms / iteration)  KLD %.4f JSD %.4f" % (
            m.name, fused, len(rows), dt, 1e3 * dt / len(rows), rows[-1]["KL"], rows[-1]["JSD"]))
