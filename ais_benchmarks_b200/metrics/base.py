"""Metric base class (reference ais_benchmarks/metrics/base.py:2-18)."""


class CMetric:
    def __init__(self):
        self.name = "base"
        self.type = "base"
        self.value = 0

    def compute(self, **kwargs):
        raise NotImplementedError

    def pre(self, **kwargs):
        raise NotImplementedError

    def post(self, **kwargs):
        raise NotImplementedError

    def reset(self, **kwargs):
        raise NotImplementedError
