import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(got, ref):
    """|got - ref| / max(1, |ref|) over the entries where the reference is finite (the tolerance definition of
    DESIGN.md: 1e-5 on log densities, 1e-4 on KLD / normalised weights); non-finite reference entries must match
    exactly."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(got[~fin]), np.isneginf(ref[~fin])), "non-finite entries differ"
    if not fin.any():
        return 0.0
    err = np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))
    worst = int(np.argmax(err))
    rel_err.last = "worst entry %d of %d: got %.9g ref %.9g" % (worst, err.size, got[fin][worst], ref[fin][worst])
    return float(err[worst])
