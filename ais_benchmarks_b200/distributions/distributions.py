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

    @loc.setter
    def loc(self, loc):
        self._loc = loc

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, scale):
        self._scale = scale

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
        """Monte-Carlo integral of distributions.py:215-226 (host RNG for the points, device for the densities)."""
        print("WARNING! Distribution specific integral not implemented for %s. Resorting to Monte-Carlo integration."
              % self.__class__.__name__)
        samples = np.random.uniform(a, b, size=(nsamples, self.dims))
        probs = np.asarray(self.prob(samples)).flatten()
        height = np.max(probs)
        points = np.random.uniform(0, height, nsamples)
        inliers_count = np.sum(probs >= points)
        volume = np.prod(np.asarray(b) - np.asarray(a)) * height
        return (inliers_count / nsamples) * volume

    def support(self):
        """(2, dims) array [lower, upper] (distributions.py:228-234)."""
        return np.array(self.support_vals).reshape(2, self.dims)

    def draw(self, ax, resolution=.01, label=None, color=None):
        # plotting is out of scope for the B200 build (DESIGN.md "Out of scope")
        return []

    def prob(self, x):
        assert len(x.shape) == 2 and x.shape[1] == self.dims, "Shape mismatch. x must be an Nx%d array. That " \
                                                             "contains N points to be evaluated." % self.dims
        if self.likelihood_f is not None:
            return self.likelihood_f(x)
        elif self.loglikelihood_f is not None:
            return np.exp(self.loglikelihood_f(x))
        raise Exception("Likelihood and LogLikelihood functions not defined")

    def log_prob(self, x):
        assert len(x.shape) == 2 and x.shape[1] == self.dims, "Shape mismatch. x must be an Nx%d array. That " \
                                                             "contains N points to be evaluated." % self.dims
        if self.loglikelihood_f is not None:
            return self.loglikelihood_f(x)
        elif self.likelihood_f is not None:
            return np.log(self.likelihood_f(x))
        raise Exception("Likelihood and LogLikelihood functions not defined")

    def is_ready(self):
        return True

    def wait_for_ready(self, timeout):
        t_ini = time.time()
        while time.time() - t_ini < timeout:
            if self.is_ready():
                return True
            time.sleep(1e-6)
        raise TimeoutError("CDistribution: wait_for_ready timed out.")

    def to_dict(self, name=None, batch_size=32, nsamples=1000, nsamples_eval=2000):
        dist_yaml = dict()
        dist_yaml["name"] = name if name is not None else self.type
        dist_yaml["type"] = "distributions." + self.__class__.__name__
        dist_yaml["batch_size"] = batch_size
        dist_yaml["nsamples"] = nsamples
        dist_yaml["nsamples_eval"] = nsamples_eval
        dist_yaml["params"] = dict()
        dist_yaml["params"]["family"] = self.family
        dist_yaml["params"]["dims"] = self.dims
        dist_yaml["params"]["support"] = [np.asarray(v).tolist() for v in self.support_vals]
        return dist_yaml
