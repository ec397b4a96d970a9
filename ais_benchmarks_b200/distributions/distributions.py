"""Base class of the B200 distributions: same contract as the reference CDistribution
(reference ais_benchmarks/distributions/distributions.py:8-307), minus the matplotlib drawing.

Batch-first: every method takes (N, dims) points and returns (N,) results, also for N == 1.
Points may be numpy arrays (results come back as numpy float64, like the reference) or torch
tensors (results stay on the GPU as float32 tensors).
"""
import time
from abc import ABCMeta, abstractmethod

import numpy as np

from .. import _device


class CDistribution(metaclass=ABCMeta):
    def __init__(self, params):
        # same required keys and assertion messages as distributions.py:87-101
        self._check_param(params, "family")
        self._check_param(params, "dims")
        self._check_param(params, "likelihood_f")
        self._check_param(params, "loglikelihood_f")
        self._check_param(params, "support")

        self.params = params
        self.type = params["type"]
        self.family = params["family"]
        self.dims = params["dims"]
        # The reference stores the bound methods self.prob / self.log_prob on the instance and in `params`
        # (distributions.py:98-99), which makes every distribution a reference cycle that only the cyclic GC can
        # free. Here a distribution can own a 64 MB packed KDE in HBM, so the default callables are resolved lazily
        # (same attribute names, same values) and the object dies with its last reference.
        lik, loglik = params["likelihood_f"], params["loglikelihood_f"]
        self._likelihood_f = None if getattr(lik, "__self__", None) is self else lik
        self._loglikelihood_f = None if getattr(loglik, "__self__", None) is self else loglik
        if self._likelihood_f is None:
            params["likelihood_f"] = "self.prob"
        if self._loglikelihood_f is None:
            params["loglikelihood_f"] = "self.log_prob"
        self.support_vals = params["support"]
        self._loc = None
        self._scale = None

    @property
    def likelihood_f(self):
        return self.prob if self._likelihood_f is None else self._likelihood_f

    @likelihood_f.setter
    def likelihood_f(self, f):
        self._likelihood_f = f

    @property
    def loglikelihood_f(self):
        return self.log_prob if self._loglikelihood_f is None else self._loglikelihood_f

    @loglikelihood_f.setter
    def loglikelihood_f(self, f):
        self._loglikelihood_f = f

    @property
    def loc(self):
        return self._loc

    def _params_changed(self):
        """Packed device parameters derived from loc / scale are stale (subclasses cache them by `_version`)."""
        self._version = getattr(self, "_version", 0) + 1
        if "_packed" in self.__dict__:          # an instance-level cache (CMultivariateNormal), not a method
            self.__dict__["_packed"] = None

    @loc.setter
    def loc(self, loc):
        self._loc = loc
        self._params_changed()

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, scale):
        self._scale = scale
        self._params_changed()

    def _check_param(self, params, param, type=None):
        assert param in params, "%s '%s' parameter is required" % (self.__class__.__name__, param)
        if type is not None:
            assert isinstance(params[param], type), "%s: Parameter '%s' is required to have type %s" % (
                self.__class__.__name__, param, str(type))

    def _check_shape(self, x):
        """distributions.py:128-147: 1-D -> (1, dims); 2-D must have dims columns; else ValueError."""
        return _device.check_points(x, self.dims)

    def set_likelihood_f(self, likelihood_f):
        self.likelihood_f = likelihood_f

    def set_loglikelihood_f(self, loglikelihood_f):
        self.loglikelihood_f = loglikelihood_f

    def sample(self, nsamples=1):
        raise NotImplementedError

    @abstractmethod
    def condition(self, dist):
        raise NotImplementedError

    @abstractmethod
    def marginal(self, dim):
        raise NotImplementedError

    def integral(self, a, b, nsamples=10000000):
        """Integral of the density over the box [a, b] (role of distributions.py:215-226, which the reference's
        test_support_integral checks to +-0.01). Mean-value Monte Carlo, box volume x E[prob(U)], evaluated in chunks of
        2^20 uniform points: the same expectation as the reference's hit-or-miss estimator with a smaller variance and
        without a 1e7-row matrix. Point draws follow np.random.seed like the reference's."""
        lo = np.asarray(a, dtype=np.float64).reshape(-1)
        hi = np.asarray(b, dtype=np.float64).reshape(-1)
        total, done = 0.0, 0
        while done < nsamples:
            m = int(min(1 << 20, nsamples - done))
            pts = lo + (hi - lo) * np.random.random_sample((m, self.dims))
            total += float(np.sum(np.asarray(self.prob(pts), dtype=np.float64)))
            done += m
        return float(np.prod(hi - lo)) * total / max(done, 1)

    def support(self):
        """(2, dims) array [lower, upper] (distributions.py:228-234)."""
        return np.array(self.support_vals).reshape(2, self.dims)

    def draw(self, ax, resolution=.01, label=None, color=None):
        # plotting is out of scope for the B200 build (DESIGN.md "Out of scope")
        return []

    def _density(self, x, want_log):
        """Generic prob / log_prob of distributions.py:251-283 for subclasses that only install a likelihood callable:
        the directly matching callable if there is one, else the other one mapped through exp / log."""
        if x.ndim != 2 or x.shape[1] != self.dims:
            raise AssertionError("Shape mismatch. x must be an Nx%d array. That contains N points to be evaluated."
                                 % self.dims)
        direct, other = ((self.loglikelihood_f, self.likelihood_f) if want_log
                         else (self.likelihood_f, self.loglikelihood_f))
        if direct is not None:
            return direct(x)
        if other is None:
            raise Exception("Likelihood and LogLikelihood functions not defined")
        return np.log(other(x)) if want_log else np.exp(other(x))

    def prob(self, x):
        return self._density(x, want_log=False)

    def log_prob(self, x):
        return self._density(x, want_log=True)

    def is_ready(self):
        return True

    def wait_for_ready(self, timeout):
        """Poll is_ready() until it holds or `timeout` seconds have passed (TimeoutError), distributions.py:288-294."""
        deadline = time.monotonic() + timeout
        while not self.is_ready():
            if time.monotonic() >= deadline:
                raise TimeoutError("CDistribution: wait_for_ready timed out.")
            time.sleep(1e-6)
        return True

    def to_dict(self, name=None, batch_size=32, nsamples=1000, nsamples_eval=2000):
        """The YAML `targets:` entry that rebuilds this distribution (schema of benchmark/def_benchmark.yaml)."""
        return {"name": self.type if name is None else name,
                "type": "distributions.%s" % type(self).__name__,
                "batch_size": batch_size, "nsamples": nsamples, "nsamples_eval": nsamples_eval,
                "params": {"family": self.family, "dims": self.dims,
                           "support": [np.asarray(v).tolist() for v in self.support_vals]}}
