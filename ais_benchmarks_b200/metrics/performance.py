"""Performance metrics with the reference's interface (reference ais_benchmarks/metrics/performance.py:15-80).

pre()/post() bracket the sampling call of the evaluation loop (benchmark/CBenchmark.py:482-490). Because the sampler
runs asynchronously on the GPU, post() synchronises the device before reading the clock, so the value is the time /
memory of the work, not of its launch.
"""
import gc
from timeit import default_timer as timer

import torch

from .base import CMetric


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


class CElapsedTime(CMetric):
    """Accumulated wall-clock seconds between pre() and post(), minus the measurement overhead (performance.py:15-52).
    The overhead is the mean of back-to-back pre()/post() pairs (the reference sleeps 10 ms inside each pair and
    subtracts the sleep; the estimate is the same quantity without the two-second start-up)."""

    def __init__(self, calibration_pairs=200):
        super().__init__()
        self.name = "time"
        self.type = "performance"
        self.t_ini = 0
        self.t_end = 0
        self.overhead = 0
        _sync()
        acc = 0.0
        for _ in range(calibration_pairs):
            self.pre()
            self.post()
            acc += self.compute()
            self.reset()
        self.overhead = acc / max(1, calibration_pairs)

    def pre(self, **kwargs):
        self.t_ini = timer()

    def post(self, **kwargs):
        _sync()
        self.t_end = timer()
        self.value += self.t_end - self.t_ini - self.overhead

    def compute(self, **kwargs):
        return self.value

    def reset(self, **kwargs):
        self.t_ini = 0
        self.t_end = 0
        self.value = 0


class CMemoryUsage(CMetric):
    """Accumulated growth, in MB, of the memory held between pre() and post() (performance.py:55-80). The reference
    measures the Python heap with guppy; the state of this package lives in HBM, so the value is the growth of the
    process' resident set plus the growth of the device memory allocated through torch."""

    def __init__(self):
        super().__init__()
        self.name = "memory"
        self.type = "performance"
        self.mem_ini = 0
        self.mem_end = 0
        self.mem_cum = 0
        import psutil
        self._proc = psutil.Process()

    def _held(self):
        gc.collect()
        dev = 0
        if torch.cuda.is_available():
            torch.cuda.synchronize()
            dev = torch.cuda.memory_allocated()
        return self._proc.memory_info().rss + dev

    def pre(self, **kwargs):
        self.mem_ini = self._held()

    def post(self, **kwargs):
        self.mem_end = self._held()
        self.mem_cum += self.mem_end - self.mem_ini
        self.value = self.mem_cum / 10 ** 6

    def compute(self, **kwargs):
        return self.value

    def reset(self, **kwargs):
        # performance.py:77-81: all three counters restart (the reported `value` is only refreshed by the next post())
        self.mem_ini = 0
        self.mem_end = 0
        self.mem_cum = 0
        gc.collect()
