"""Evaluation loop of the reference (ais_benchmarks/benchmark/CBenchmark.py) with the state kept in HBM."""
from .CBenchmark import CBenchmark

__all__ = ["CBenchmark"]
