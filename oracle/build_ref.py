"""Recipe that makes the UNMODIFIED reference travel to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY. The reference (IntelLabs/ais-benchmarks) is pure Python: there is nothing to
compile. "Building" it means copying, byte for byte, the package and its own unit tests from where they lie under
/root/reference into the git-ignored directory `oracle/_ref/` (listed in .gitignore, NOT in .gpurunignore, so it
ships with `gpurun` exactly like a built .so, and never enters the history). Nothing is edited; `oracle/_ref/MANIFEST`
records the sha256 of every copied file so that the copy can be audited against the source tree.

    python oracle/build_ref.py            # run here (the build container); __graft_entry__.build() calls it too

Consumers (the only ones allowed to touch oracle/):
  * tests/test_reference_dropin.py      the reference's own unittest files and CBenchmark.evaluate_method run on top of
                                        this repo's classes
  * bench.py --impl reference / cpu_baseline   time the reference's CKLDivergence / CMixtureModel / importance_sample
The product package (ais_benchmarks_b200/) never imports it.

Import recipe: stub matplotlib / mpl_toolkits / guppy in sys.modules, add np.int, then put
oracle/_ref on sys.path -- `load()` below does that and returns the reference's top-level package.
"""
import hashlib
import os
import shutil
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("AIS_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")
# the whole package (464 KB of .py / .yaml: sampling_methods/tree_pyramid.py imports visualization/ at module level,
# so nothing can be left out), the reference's own tests and its licence
COPY = ["ais_benchmarks", "tests", "LICENSE"]
SKIP_DIRS = {"__pycache__"}


def _files(root):
    for base, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if not f.endswith((".pyc", ".pyo")):
                yield os.path.join(base, f)


def build(verbose=True):
    """Copy the reference files into oracle/_ref/ (idempotent). Returns the number of files, 0 if the source tree is
    absent (GPU box: the prebuilt copy is used as is)."""
    if not os.path.isdir(os.path.join(REF_SRC, "ais_benchmarks")):
        return 0
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    manifest = []
    for rel in COPY:
        src = os.path.join(REF_SRC, rel)
        srcs = [src] if os.path.isfile(src) else list(_files(src))
        for s in srcs:
            r = os.path.relpath(s, REF_SRC)
            d = os.path.join(REF_DST, r)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            with open(s, "rb") as f:
                manifest.append("%s  %s" % (hashlib.sha256(f.read()).hexdigest(), r))
    with open(os.path.join(REF_DST, "MANIFEST"), "w") as f:
        f.write("# sha256 of files copied verbatim from %s by oracle/build_ref.py\n" % REF_SRC)
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("oracle/_ref: %d reference files copied from %s" % (len(manifest), REF_SRC))
    return len(manifest)


def available():
    return os.path.isfile(os.path.join(REF_DST, "ais_benchmarks", "distributions", "__init__.py"))


def verify():
    """Check the copy against its manifest (True / False)."""
    try:
        with open(os.path.join(REF_DST, "MANIFEST")) as f:
            lines = [ln.split("  ", 1) for ln in f.read().splitlines() if ln and not ln.startswith("#")]
        for sha, rel in lines:
            with open(os.path.join(REF_DST, rel), "rb") as f:
                if hashlib.sha256(f.read()).hexdigest() != sha:
                    return False
        return len(lines) > 0
    except OSError:
        return False


_STUBS = ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches", "matplotlib.lines",
          "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d", "guppy"]


def load():
    """-> the reference's top-level package `ais_benchmarks`, imported from oracle/_ref (unmodified files)."""
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # sampling_methods/tree_pyramid.py:395 only
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import ais_benchmarks  # noqa: F401
    if not os.path.abspath(ais_benchmarks.__path__[0]).startswith(REF_DST):
        raise ImportError("another `ais_benchmarks` package shadows oracle/_ref: %s" % ais_benchmarks.__path__[0])
    import ais_benchmarks.distributions  # noqa: F401
    import ais_benchmarks.sampling_methods  # noqa: F401
    import ais_benchmarks.metrics.divergences  # noqa: F401
    return sys.modules["ais_benchmarks"]


def seed_tests_namespace():
    """The venv ships a regular package called `tests` that shadows the reference's namespace package of the same
    name; the reference's test files import `tests.distributions.test_distribution_template`.
    Pre-seed sys.modules with namespace modules that point into oracle/_ref/tests. Returns the modules it replaced so
    that the caller can restore them."""
    saved = {}
    for mod, sub in (("tests", "tests"), ("tests.distributions", "tests/distributions"), ("tests.metrics", "tests/metrics")):
        saved[mod] = sys.modules.get(mod)
        m = types.ModuleType(mod)
        m.__path__ = [os.path.join(REF_DST, sub)]
        sys.modules[mod] = m
    return saved


if __name__ == "__main__":
    n = build()
    if n == 0:
        print("reference tree not found at %s; nothing copied" % REF_SRC)
        sys.exit(1)
    assert verify()
