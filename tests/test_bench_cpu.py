"""bench.py pieces that need no GPU: the synthetic workload generators, the CPU arm (the UNMODIFIED reference from
oracle/_ref when it has been built, else the oracle port of the same algorithm) and the JSON contract of
`bench.py --impl reference`."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ais_oracle as O  # noqa: E402


def test_workload_generators_are_seeded_and_shaped():
    mu, sig, w, sup, centres, x = bench.kde_workload(scale=0.002)
    assert centres.shape == (2000, 8) and x.shape == (2000, 8) and centres.dtype == np.float32
    assert mu.shape == (16, 8) and sig.shape == (16, 8) and abs(w.sum() - 1.0) < 1e-12
    assert np.all(sig >= 0.04) and np.all(sig <= 0.05)                    # variances, as generateRandomGMM
    assert np.all(x >= sup[0] - 1e-6) and np.all(x <= sup[1] + 1e-6)      # evaluation points ~ U(support)
    again = bench.kde_workload(scale=0.002)
    assert np.array_equal(again[4], centres) and np.array_equal(again[5], x)
    mu2, sig2, w2, sup2, pmeans, per = bench.dm_workload(scale=0.001)
    assert mu2.shape == (64, 32) and pmeans.shape == (1024, 32) and per == 9
    assert np.all(pmeans >= sup2[0]) and np.all(pmeans <= sup2[1])


def test_cpu_arm_matches_a_direct_evaluation():
    dt, kld, kind = bench.cpu_kde_once(64, 300, 1, kind="port")
    assert dt > 0 and np.isfinite(kld) and kind == "port"
    # same numbers from the oracle's closed forms on the same seeded inputs
    c = bench.KDE_CFG
    mu, sig, w, sup = bench.random_gmm(0, c["D"], c["modes"])
    rs = np.random.RandomState(1)
    centres = bench.gmm_samples(rs, mu, sig, w, 300).astype(np.float64)
    x = rs.uniform(sup[0], sup[1], size=(64, c["D"]))
    ref = O.kld_from_log_probs(O.gmm_log_prob(x, mu, sig, w), O.iso_mixture_log_prob(x, centres, c["bw"]))
    assert abs(kld - ref) <= 1e-9 * max(1.0, abs(ref))
    dt2, kld2, _ = bench.cpu_kde_once(64, 300, 2, kind="port")            # two worker processes: same value
    assert abs(kld2 - kld) <= 1e-9 * max(1.0, abs(kld))


def test_reference_cpu_arm_runs_the_real_reference():
    """With oracle/_ref present (oracle/build_ref.py) the CPU arm drives the reference's own classes. Its KDE resamples
    the centres with replacement (base.py:164, seeded through np.random.seed), so the value is compared with the oracle
    evaluated on the SAME resampled centres."""
    from oracle import build_ref
    if not build_ref.available() and build_ref.build(verbose=False) == 0:
        import pytest
        pytest.skip("oracle/_ref not built and /root/reference absent")
    dt, kld, kind = bench.cpu_kde_once(64, 300, 1, kind="reference")
    assert kind == "reference" and np.isfinite(kld)
    q = bench._REF_STATE["q"]
    centres = np.array([m.loc for m in q.model.models])
    c = bench.KDE_CFG
    mu, sig, w, sup = bench.random_gmm(0, c["D"], c["modes"])
    rs = np.random.RandomState(1)
    bench.gmm_samples(rs, mu, sig, w, 300)
    x = rs.uniform(sup[0], sup[1], size=(64, c["D"]))
    ref = O.kld_from_log_probs(O.gmm_log_prob(x, mu, sig, w), O.iso_mixture_log_prob(x, centres, c["bw"]))
    assert abs(kld - ref) <= 1e-9 * max(1.0, abs(ref))
    dt2, kld2, _ = bench.cpu_kde_once(64, 300, 2, kind="reference")       # forked workers share the same KDE
    assert abs(kld2 - kld) <= 1e-9 * max(1.0, abs(kld))


def test_reference_arm_json_contract():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                                # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0
    cb = d["cpu_baseline"]
    from oracle import build_ref
    assert cb["kind"] == ("reference" if build_ref.available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    if cb["kind"] == "reference":                                         # DM-AIS as shipped + batched
        dm = d["dm_ais"]["cpu_baseline"]
        assert dm["kind"] == "reference" and dm["value"] > 0 and dm["batched"]["value"] > dm["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    # non-zero ranks of a torchrun launch print nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
