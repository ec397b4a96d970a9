"""Drop-in demonstration ("What's missing" 1 / W3): the reference's OWN unit tests and its OWN caller
(`CBenchmark.load_benchmark/load_methods/evaluate_method/write_results`) run, unmodified, on top of this repo's classes.

The reference ships to the GPU box as `oracle/_ref/` (oracle/build_ref.py: a verbatim, checksummed copy of
/root/reference, git-ignored). The work happens in tests/ref_dropin_runner.py, in a subprocess, because it rewires
`sys.modules` (module aliasing per INTEGRATION.md section A and the `tests` namespace of SURVEY.md section 4).
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def _ensure_ref():
    if not build_ref.available():
        build_ref.build(verbose=False)      # only possible where /root/reference exists (the build container)
    if not build_ref.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")


def _run(*args, timeout=900):
    proc = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "tests", "ref_dropin_runner.py"), *args],
                          cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("DROPIN_JSON ")]
    assert proc.returncode == 0 and lines, "runner failed:\n%s\n%s" % (proc.stdout[-3000:], proc.stderr[-3000:])
    return json.loads(lines[-1][len("DROPIN_JSON "):])


def test_reference_copy_is_verbatim():
    """oracle/_ref matches its sha256 manifest, and (where the source tree is present) the source tree itself."""
    _ensure_ref()
    assert build_ref.verify()
    if os.path.isdir(os.path.join(build_ref.REF_SRC, "ais_benchmarks")):
        import hashlib
        with open(os.path.join(build_ref.REF_DST, "MANIFEST")) as f:
            entries = [ln.split("  ", 1) for ln in f.read().splitlines() if ln and not ln.startswith("#")]
        assert len(entries) > 40
        for sha, rel in entries:
            with open(os.path.join(build_ref.REF_SRC, rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == sha, rel


def test_harness_runs_the_reference_on_itself():
    """CPU leg: the same runner with the reference's own classes -- 27 of the reference's unit tests pass (SURVEY.md
    section 4 counts 21 + the 6 pre/post cases) -- so a failure of the GPU leg below is a failure of OUR classes."""
    _ensure_ref()
    out = _run("--impl", "reference", "--skip-benchmark")
    assert out["ref_verified"] and out["unittest"]["run"] >= 21
    assert not out["unittest"]["failures"] and not out["unittest"]["errors"], out["unittest"]


@pytest.mark.gpu
def test_reference_tests_and_cbenchmark_run_on_b200_classes():
    """GPU leg: classes aliased per INTEGRATION.md section A; the reference's MVN / Uniform / KLD unittest files pass and
    the reference's CBenchmark evaluates DM-AIS, M-PMC, LAIS and TP-AIS on the default targets through our objects."""
    _ensure_ref()
    out = _run("--impl", "b200")
    assert out["ref_verified"]
    ut = out["unittest"]
    assert ut["run"] >= 21 and not ut["failures"] and not ut["errors"], ut
    bm = out["benchmark"]
    assert bm["experiments"] == 12, bm                      # 3 targets x 4 samplers
    for key, row in bm["final"].items():
        assert row["KLD"] >= -1e-6 and row["JSD"] >= -1e-6, (key, row)
    # the adaptive samplers end close to the target on the easy 1-D Normal (reference run: 0.06 .. 0.25)
    assert bm["final"]["TP_AISr_ESS90/normal01D"]["KLD"] < 0.5
    assert bm["final"]["dm_ais/normal01D"]["KLD"] < 1.0


def test_reference_tpais_dm_and_mixture_methods_cannot_run():
    """Why CTreePyramidSampling here offers method="simple" only (VERDICT r1 "missing" 5): with the
    `np.int` alias the reference's own "dm" / "mixture" weighting (tree_pyramid.py:363-419) still dies -- the first
    `find_unique` after the root has been split recurses without bound in `CTreePyramid.find` (:140-155) -- for both
    kernels. There is no reference behaviour to reproduce; ours raises NotImplementedError and says so."""
    _ensure_ref()
    probe = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import build_ref
build_ref.load()
np.int = int
from ais_benchmarks.sampling_methods import CTreePyramidSampling
from ais_benchmarks.distributions import CGaussianMixtureModel
tgt = CGaussianMixtureModel({"means": np.array([[0.2, 0.3], [-0.4, -0.1]]), "sigmas": np.array([[0.02, 0.02], [0.03, 0.03]]),
                             "weights": np.array([0.5, 0.5]), "support": np.array([[-1., -1.], [1., 1.]])})
for method in ("mixture", "dm", "simple"):
    for kernel in ("haar", "normal"):
        np.random.seed(0)
        try:
            s = CTreePyramidSampling({"space_min": np.array([-1., -1.]), "space_max": np.array([1., 1.]), "dims": 2,
                                      "method": method, "resampling": "none", "kernel": kernel, "ess_target": 0.5,
                                      "n_min": 5, "parallel_samples": 4})
            x, w = s.importance_sample(tgt, 40, timeout=20)
            print("PROBE", method, kernel, "ok", len(x))
        except BaseException as e:
            print("PROBE", method, kernel, type(e).__name__)
''' % ROOT
    proc = subprocess.run([sys.executable, "-W", "ignore", "-c", probe], cwd=ROOT, capture_output=True, text=True,
                          timeout=600)
    got = {tuple(ln.split()[1:3]): ln.split()[3] for ln in proc.stdout.splitlines() if ln.startswith("PROBE ")}
    assert len(got) == 6, (proc.stdout[-2000:], proc.stderr[-2000:])
    for kernel in ("haar", "normal"):
        assert got[("simple", kernel)] == "ok"
        assert got[("mixture", kernel)] == "RecursionError" and got[("dm", kernel)] == "RecursionError", got
    # ours: refuses at construction, with the explanation
    from ais_benchmarks_b200.sampling_methods import CTreePyramidSampling
    import numpy as np
    with pytest.raises(NotImplementedError, match="np.int"):
        CTreePyramidSampling({"space_min": np.array([-1., -1.]), "space_max": np.array([1., 1.]), "dims": 2,
                              "method": "mixture", "resampling": "none", "kernel": "haar", "ess_target": 0.5,
                              "n_min": 5, "parallel_samples": 4})
