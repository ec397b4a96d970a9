"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares, and its
host-side functions (packing, partial-record algebra) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
from kernel_model import packed_logpdf
from oracle import ais_oracle as O

from ais_benchmarks_b200 import _lib
from ais_benchmarks_b200._device import DeviceMixture, kld_finalize, kld_merge, weight_ess, weight_logsum, weight_merge


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ais_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aisb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libais_b200.so does not export %s" % s
    assert set(syms) == set(_lib.PROTOTYPES), set(syms) ^ set(_lib.PROTOTYPES)
    assert lib.aisb_abi_version() == 1


def test_layout_queries():
    lib = _lib.load()
    assert [lib.aisb_padded_dims(d) for d in (1, 2, 3, 4, 5, 8, 9, 32, 33, 64, 65, 0)] == \
        [1, 2, 4, 4, 8, 8, 16, 32, 64, 64, 0, 0]
    assert lib.aisb_padded_components(1) == 32 and lib.aisb_padded_components(33) == 64
    assert lib.aisb_packed_row_floats(0, 8) == 8 and lib.aisb_packed_row_floats(1, 8) == 16
    assert lib.aisb_packed_row_floats(2, 8) == 44 and lib.aisb_packed_row_floats(2, 3) == 16
    assert lib.aisb_packed_row_floats(2, 64) == 0
    assert lib.aisb_iso_scale(0.02) == pytest.approx(np.sqrt(np.log2(np.e) / 0.04), rel=1e-6)


@pytest.mark.parametrize("case", range(6))
def test_pack_diag_matches_oracle(case):
    g = load_golden("gmm")
    means, sigmas, w, x = g["means%d" % case], g["sigmas%d" % case], g["weights%d" % case], g["x%d" % case]
    K, D = means.shape
    rows, offs, _ = DeviceMixture.pack_host(means, sigmas, w, _lib.COV_DIAG)
    DP = _lib.load().aisb_padded_dims(D)
    assert rows.shape == (_lib.load().aisb_padded_components(K), 2 * DP)
    assert np.all(np.isneginf(offs[K:])) and np.all(rows[K:] == 0)
    got = packed_logpdf(x, rows, offs, K, D, 1, 0.0, DP)
    ref = g["logp%d" % case]
    assert np.max(np.abs(got - ref) / np.maximum(1, np.abs(ref))) < 2e-6  # float32 parameter rounding only


def test_pack_full_and_weights_match_oracle():
    g = load_golden("mixture")
    means, covs, w, x = g["means"], g["covs"], g["weights"], g["x"]
    K, D = means.shape
    rows, offs, _ = DeviceMixture.pack_host(means, covs, w.copy(), _lib.COV_FULL)
    assert np.isneginf(offs[[1, 2, 4]]).all()  # zero / NaN / negative weights
    got = packed_logpdf(x, rows, offs, K, D, 2, 0.0, 4)
    assert np.max(np.abs(got - g["logp"]) / np.maximum(1, np.abs(g["logp"]))) < 2e-6


def test_pack_full_nonsymmetric_cov_matches_reference():
    g = load_golden("mvn")
    i = 4  # CMultivariateNormal.py __main__ demo covariance [[0.1, 0.2], [0, 0.2]]
    rows, offs, _ = DeviceMixture.pack_host(g["mean%d" % i].reshape(1, -1), g["cov%d" % i].reshape(1, 2, 2), None,
                                            _lib.COV_FULL)
    got = packed_logpdf(g["x%d" % i], rows, offs, 1, 2, 2, 0.0, 2)
    assert np.max(np.abs(got - g["logp%d" % i]) / np.maximum(1, np.abs(g["logp%d" % i]))) < 2e-6


def test_pack_iso_matches_oracle():
    g = load_golden("dm_weights")
    pm, x, sigma = g["pmeans0"], g["x0"], float(g["sigma0"])
    rows, offs, iso = DeviceMixture.pack_host(pm, np.array([sigma]), None, _lib.COV_ISO_SHARED)
    got = packed_logpdf(x, rows, offs, pm.shape[0], 2, 0, iso, 2)
    assert np.max(np.abs(got - g["logq0"]) / np.maximum(1, np.abs(g["logq0"]))) < 2e-6


def test_pack_rejects_bad_covariance():
    with pytest.raises(AssertionError):
        DeviceMixture.pack_host(np.zeros((1, 2)), np.array([[1.0, -0.5]]), None, _lib.COV_DIAG)
    with pytest.raises(AssertionError):
        DeviceMixture.pack_host(np.zeros((1, 2)), np.array([[[1.0, 2.0], [2.0, 1.0]]]), None, _lib.COV_FULL)  # det < 0


def test_weight_partial_algebra():
    rs = np.random.RandomState(0)
    lw = rs.normal(size=1000) * 5

    def record(v):
        m = v.max()
        return np.array([m, np.exp(v - m).sum(), np.exp(2 * (v - m)).sum(), len(v)])

    merged = weight_merge(record(lw[:300]), record(lw[300:]))
    w = np.exp(lw)
    assert weight_logsum(merged) == pytest.approx(np.log(w.sum()), rel=1e-12)
    assert weight_ess(merged) / 1000 == pytest.approx(O.approx_ness(w), rel=1e-12)
    assert merged[3] == 1000


def test_kld_partial_algebra():
    g = load_golden("kde_kld")
    lp, lq = g["lp1"], g["lq1"]

    def record(p, q):
        inl = (p > -np.inf) & (q > -np.inf)
        p, q = p[inl], q[inl]
        mp, mq = (p.max(), q.max()) if len(p) else (-3e38, -3e38)
        return np.array([mp, np.exp(p - mp).sum(), (np.exp(p - mp) * (p - q)).sum(), mq, np.exp(q - mq).sum(),
                         inl.sum(), 0, len(inl)])

    half = len(lp) // 3
    merged = kld_merge(record(lp[:half], lq[:half]), record(lp[half:], lq[half:]))
    assert kld_finalize(merged) == pytest.approx(float(g["kld1"]), rel=1e-10)
    assert kld_finalize(record(g["edge_lp"], g["edge_lq"])) == pytest.approx(float(g["edge_kld"]), rel=1e-12)
    assert kld_finalize(record(np.full(4, -np.inf), np.zeros(4))) == 0.0
    inf_rec = record(lp, lq)
    inf_rec[6] = 1
    assert kld_finalize(inf_rec) == np.inf
