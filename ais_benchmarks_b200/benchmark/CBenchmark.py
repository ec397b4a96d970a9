"""The caller of the hot path: the (target, method) evaluation loop of the reference
(ais_benchmarks/benchmark/CBenchmark.py:58-137 loaders, :151-199 run, :318-326 write_results, :328-544
evaluate_method) with the same YAML schema, the same results.txt format and the same method signatures -- and the loop
itself FUSED around device-resident state:

  reference, every iteration                      here, every iteration
  ------------------------------------------     ---------------------------------------------------------
  importance_sample -> numpy samples, weights     importance_sample with device outputs (nothing leaves HBM)
  for each metric:                                one evaluation set x (drawn once per experiment) and its
      draw nsamples_eval uniform points               target log density lp (evaluated once per experiment)
      p.log_prob / p.prob on them                 ONE q evaluation  lq = q.log_prob(x)   (the pairwise kernel)
      q.log_prob / q.prob on them                 KLD, JSD, EVMSE all reduced from (x, lp, lq) on the device:
      numpy reductions                                only 8 + 8 + 2 + 2 x (4 + D) doubles reach the host
  write one results.txt row                       write one results.txt row (same text format)

`fused=False` restores the reference's data flow (every metric draws its own points through np.random and evaluates
both densities itself). Plotting, animation and the benchden targets are out of scope; the
`debug` / plot arguments are accepted and ignored.
"""
import ast
import os
import pathlib
import sys
import time

import numpy as np
import torch
import yaml

from .. import _device
from .. import distributions as _distributions
from .. import sampling_methods as _sampling_methods
from ..metrics import CElapsedTime, CExpectedValueMSE, CJSDivergence, CKLDivergence, CMemoryUsage
from ..metrics.divergences import _log_prob_device, evaluation_points


def time_to_hms(seconds):
    """utils/misc.py time_to_hms: (hours, minutes, seconds)."""
    minutes, sec = divmod(seconds, 60)
    hours, minutes = divmod(minutes, 60)
    return int(hours), int(minutes), sec


def _registry(module):
    return {name: getattr(module, name) for name in module.__all__}


def _literal(value):
    """A YAML parameter value as the reference's eval() would see it (CBenchmark.py:68-70): numbers pass through,
    strings are Python literals ("'simple'" -> 'simple'); anything that is not a literal stays a string."""
    if not isinstance(value, str):
        return value
    try:
        return ast.literal_eval(value)
    except (ValueError, SyntaxError):
        return value


class CBenchmark(object):
    # classes the YAML files may name (the reference resolves them with eval() in a star-import namespace, :15-16)
    TARGETS = _registry(_distributions)
    METHODS = _registry(_sampling_methods)
    METRICS = {"KLD": CKLDivergence, "JSD": CJSDivergence, "EVMSE": CExpectedValueMSE, "T": CElapsedTime,
               "MEM": CMemoryUsage}

    def __init__(self):
        self.targets = []        # target distributions
        self.ndims = []          # their dimensionality
        self.methods = []        # sampling methods under evaluation
        self.batch_sizes = []    # samples requested per iteration, per target
        self.nsamples = []       # samples after which a (method, target) evaluation ends, per target
        self.timeout = 18000     # seconds allowed per (method, target)
        self.eval_sampl = []     # evaluation points of the metrics, per target
        self.metrics = []        # metric names
        self.n_experiments = 10  # repetitions of every (method, target)
        self.rseed = 0
        self.output_file = "results.txt"
        self.debug_text = True
        self.fused = True        # see module docstring
        self.skipped_methods = []

    # -- loaders (same YAML schema as benchmark/def_*.yaml) ------------------------------------------------------
    def load_config(self, config_file):
        with open(config_file, mode="r") as f:
            cfg = yaml.load(f, Loader=yaml.SafeLoader)
        self.metrics = list(cfg["metrics"])
        self.debug_text = bool(cfg.get("debug", {}).get("text", True))
        self.n_experiments = cfg["nreps"]
        self.rseed = cfg["rseed"]
        if self.rseed >= 0:
            np.random.seed(self.rseed)
            if torch.cuda.is_available():
                _device.seed(self.rseed)

    def load_benchmark(self, benchmark_file):
        for lst in (self.targets, self.ndims, self.batch_sizes, self.eval_sampl, self.nsamples):
            lst.clear()
        with open(benchmark_file, mode="r") as f:
            bench = yaml.load(f, Loader=yaml.SafeLoader)
        for target in bench["targets"]:
            try:
                target_dist = self.TARGETS[target["type"]](dict(target["params"]))
            except BaseException as e:
                print(e)
                raise ValueError("Error creating target dist: %s(%s)" % (target["type"], target["params"]))
            target_dist.name = target["name"]
            self.nsamples.append(target["nsamples"])
            self.eval_sampl.append(target["nsamples_eval"])
            self.batch_sizes.append(target["batch_size"])
            self.targets.append(target_dist)
            self.ndims.append(target_dist.dims)

    def load_methods(self, methods_file, space_min, space_max, dims):
        """Instantiate every method of the file for one target space (CBenchmark.py:58-80). Methods this package does
        not implement (they are not on the accelerated path) are skipped and listed in self.skipped_methods."""
        self.methods.clear()
        self.skipped_methods.clear()
        with open(methods_file, mode="r") as f:
            methods = yaml.load(f, Loader=yaml.SafeLoader)
        for method in methods["methods"]:
            cls = self.METHODS.get(method["type"])
            if cls is None:
                self.skipped_methods.append(method["name"])
                print("CBenchmark: method %s (%s) is not implemented by ais_benchmarks_b200, skipped"
                      % (method["name"], method["type"]), file=sys.stderr)
                continue
            params = {"space_min": np.array(space_min, dtype=np.float64),
                      "space_max": np.array(space_max, dtype=np.float64), "dims": dims}
            for key, value in (method.get("params") or {}).items():
                params[key] = _literal(value)
            m = cls(params=params)
            m.name = method["name"]
            m.debug = method.get("debug", False)
            self.methods.append(m)

    # -- results.txt ---------------------------------------------------------------------------------------------
    @staticmethod
    def results_header(metrics):
        return ("dims output_samples " + " ".join([m for m in metrics]) +
                " NESS method target_d accept_rate proposal_samples proposal_evals target_evals\n")

    @staticmethod
    def write_results(results, method, target, nsamples, file):
        """One row per iteration, in the reference's text format (CBenchmark.py:318-326)."""
        with open(file, mode="a+") as f:
            text = "%02d %04d " % (target.dims, nsamples)
            text += " ".join(["%9.5f" % val for val in results.values()])
            text += "%9.3f %s %s %5.3f %d %d %d" % (method.get_NESS(), method.name, target.name,
                                                    method.get_acceptance_rate(), method.num_proposal_samples,
                                                    method.num_proposal_evals, method.num_target_evals)
            f.write(text + "\n")

    @staticmethod
    def read_results(file):
        """results.txt back as {column: list}; numeric columns as float (the table plot_results.py reads with pandas)."""
        with open(file) as f:
            cols = f.readline().split()
            # the reference's row format glues the last metric to NESS when the value is wide; split on whitespace only
            rows = [line.split() for line in f if line.strip()]
        table = {c: [] for c in cols}
        for r in rows:
            if len(r) != len(cols):
                raise ValueError("results row has %d fields, header has %d: %r" % (len(r), len(cols), r))
            for c, v in zip(cols, r):
                try:
                    table[c].append(float(v))
                except ValueError:
                    table[c].append(v)
        return table

    # -- run -----------------------------------------------------------------------------------------------------
    def run(self, benchmark_file, methods_file, config_file, out_file):
        self.output_file = out_file
        self.load_config(config_file)
        self.load_benchmark(benchmark_file)
        assert len(self.targets) > 0
        assert len(self.targets) == len(self.ndims) == len(self.nsamples)
        parent = pathlib.Path(self.output_file).parent
        if not parent.exists():
            os.makedirs(parent)
        with open(self.output_file, "w") as f:
            f.write(self.results_header(self.metrics))
        t_start = time.time()
        for target_dist, ndims, max_samples, eval_sampl, batch_size in \
                zip(self.targets, self.ndims, self.nsamples, self.eval_sampl, self.batch_sizes):
            self.load_methods(methods_file, target_dist.support()[0], target_dist.support()[1], ndims)
            for sampling_method in self.methods:
                print("EVALUATING: %s || dims: %d || max samples: %d || target_d: %s || batch: %d" % (
                    sampling_method.name, ndims, max_samples, target_dist.name, batch_size))
                sampling_method.reset()
                t_ini = time.time()
                CBenchmark.evaluate_method(ndims=ndims, target_dist=target_dist, sampling_method=sampling_method,
                                           max_samples=max_samples, max_sampling_time=self.timeout,
                                           batch_size=batch_size, debug=False, metrics=self.metrics, rseed=self.rseed,
                                           n_reps=self.n_experiments, sampling_eval_samples=eval_sampl,
                                           filename=self.output_file, fused=self.fused, debug_text=self.debug_text)
                print("TOOK: %dh %dm %4.1fs" % time_to_hms(time.time() - t_ini))
        print("BENCHMARK TOOK: %dh %dm %4.1fs" % time_to_hms(time.time() - t_start))

    # -- the evaluation loop ---------------------------------------------------------------------------------------
    @staticmethod
    def evaluate_method(ndims, target_dist, sampling_method, max_samples, sampling_eval_samples,
                        metrics=("NESS", "JSD", "T"), rseed=-1, n_reps=10, batch_size=16,
                        debug=True, filename=None, max_sampling_time=600, profile=False, debug_plot_path=None,
                        fused=True, debug_text=False):
        """CBenchmark.py:328-544 (arguments and their meaning unchanged; `debug`, `profile`, `debug_plot_path` are
        accepted for signature compatibility and ignored). Returns the rows written, one dict per iteration."""
        if rseed >= 0:
            np.random.seed(rseed)
            _device.seed(rseed)
        metrics_eval = [CBenchmark.METRICS[m]() for m in metrics if m in CBenchmark.METRICS]  # "NESS" is a column of
        density_metrics = [m for m in metrics_eval if m.type in ("divergence", "statistics")]  # every row (:420-421)
        rows = []
        keep_outputs = sampling_method.device_outputs
        sampling_method.device_outputs = True  # samples / weights stay in HBM between the iterations
        try:
            for nexp in range(n_reps):
                t_start = time.time()
                sampling_time = 0
                sampling_method.reset()
                n_have = 0
                n_samples = batch_size
                [m.reset() for m in metrics_eval]
                x_eval = lp_eval = None
                if fused and density_metrics:
                    # one evaluation set and one target evaluation per experiment (the reference redraws the points and
                    # re-evaluates p for every metric in every iteration, divergences.py:84-87)
                    pts = evaluation_points(density_metrics[0], {"p": target_dist, "q": sampling_method,
                                                                 "nsamples": sampling_eval_samples})
                    x_eval, was_numpy = _device.to_device_points(pts, target_dist.dims)
                    lp_eval = _log_prob_device(target_dist, x_eval, lambda: pts if was_numpy else x_eval.cpu().numpy())
                while n_have < max_samples:
                    if debug_text:
                        hr, mn, sec = time_to_hms(time.time() - t_start)
                        print("%02d/%02d | %12s | %6s | %dD | #s: %.1fk | %5.1f%% | t: %02dh %02dm %4.1fs | " % (
                            nexp + 1, n_reps, sampling_method.name, target_dist.name, ndims, n_have / 1000.0,
                            (n_have / max_samples) * 100, hr, mn, sec) +
                            " | ".join(["%s: %7.5f" % (m.name, m.value) for m in metrics_eval]))
                    [m.pre() for m in metrics_eval]
                    t0 = time.time()
                    samples_acc, _ = sampling_method.importance_sample(target_d=target_dist, n_samples=n_samples,
                                                                       timeout=max_sampling_time - sampling_time)
                    [m.post() for m in metrics_eval]
                    sampling_time += time.time() - t0
                    n_have = len(samples_acc)
                    n_samples = n_have + batch_size  # the samplers are cumulative (:489-493)

                    results = dict()
                    if fused and density_metrics:
                        lq_eval = _log_prob_device(sampling_method, x_eval, lambda: x_eval.cpu().numpy().astype(np.float64))
                    for m in metrics_eval:
                        if m not in density_metrics:
                            results[m.name] = m.compute(p=target_dist, q=sampling_method, nsamples=sampling_eval_samples)
                        elif not fused:
                            results[m.name] = m.compute(p=target_dist, q=sampling_method, nsamples=sampling_eval_samples)
                        elif isinstance(m, CExpectedValueMSE):
                            results[m.name] = m.compute_from_device_log_probs(x_eval, lp_eval, lq_eval)
                        else:
                            results[m.name] = m.compute_from_device_log_probs(lp_eval, lq_eval)
                    rows.append(dict(results, output_samples=n_have, experiment=nexp))
                    if filename is not None:
                        CBenchmark.write_results(results=results, method=sampling_method, target=target_dist,
                                                 nsamples=n_have, file=filename)
                    if sampling_time > max_sampling_time:
                        break
        finally:
            sampling_method.device_outputs = keep_outputs
        return rows
