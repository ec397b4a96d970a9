"""Import the UNMODIFIED reference (/root/reference) in this container.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py to produce the committed fixtures and by
the optional side-by-side tests that run only where /root/reference exists). The reference needs
matplotlib / guppy at import time for plotting and heap metrics; nothing numeric touches them, so they
are stubbed in sys.modules.
"""
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("AIS_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ais_benchmarks"))


def load():
    """-> the reference's top-level package `ais_benchmarks` (imported once)."""
    if not available():
        raise ImportError("reference tree not found at %s" % REFERENCE_ROOT)
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches", "matplotlib.lines",
                 "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d", "guppy"]:
        sys.modules.setdefault(name, MagicMock())
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # sampling_methods/tree_pyramid.py:395 only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import ais_benchmarks  # noqa: F401
    import ais_benchmarks.distributions  # noqa: F401
    import ais_benchmarks.sampling_methods  # noqa: F401
    import ais_benchmarks.metrics.divergences  # noqa: F401
    return sys.modules["ais_benchmarks"]
