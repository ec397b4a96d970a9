"""Expected-value MSE metric with the reference's interface (reference ais_benchmarks/metrics/statistics.py:11-48).

    EV_P = sum_i x_i p(x_i) / sum_i p(x_i),   EV_Q likewise,   EVMSE = || EV_P - EV_Q ||^2

on uniform evaluation points x_i. Here the densities stay in the log domain on the device: one streaming pass per
side for the normaliser (aisb_logw_partials_f32) and one for the weighted coordinate sums (aisb_weighted_mean_f32,
float64 accumulation). The evaluation points and both log-density vectors are the ones the KLD / JSD use, so the fused
evaluation loop computes them once for all three metrics.
"""
import numpy as np

from .. import _device
from .base import CMetric
from .divergences import device_log_probs, evaluation_points


class CExpectedValueMSE(CMetric):
    def __init__(self):
        super().__init__()
        self.name = "EVMSE"
        self.type = "statistics"
        self.device_points = False
        self.sharded = False
        self.ev_p = None
        self.ev_q = None

    def pre(self, **kwargs):
        pass

    def post(self, **kwargs):
        pass

    def reset(self, **kwargs):
        pass

    def compute(self, **kwargs):
        return self.compute_from_samples(kwargs.get("p"), kwargs.get("q"), evaluation_points(self, kwargs))

    def compute_from_samples(self, p, q, samples):
        x, lp, lq = device_log_probs(p, q, samples)
        return self.compute_from_device_log_probs(x, lp, lq)

    def _expectation(self, x, logw):
        from .. import parallel
        part = _device.logw_partials(logw).cpu().numpy()
        if self.sharded:
            part = parallel.merge_weight_partials(parallel.all_gather_records(part))
        with np.errstate(all="ignore"):
            log_norm = _device.weight_logsum(part) if part[3] > 0 else -np.inf
        ev = _device.weighted_mean(x, logw, log_norm)
        return parallel.all_gather_records(ev).sum(axis=0) if self.sharded else ev.cpu().numpy()

    def compute_from_device_log_probs(self, x, lp, lq):
        self.ev_p = self._expectation(x, lp)
        self.ev_q = self._expectation(x, lq)
        d = self.ev_p - self.ev_q
        self.value = np.float64(d @ d)
        return self.value
