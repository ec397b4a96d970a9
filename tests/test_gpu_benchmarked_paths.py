"""Parity of the code paths that bench.py TIMES, at the sizes it times them:

  * BASELINE configs[4] through the hinted tensor-core path (aisb_dm_weights_hinted_f32 -> tc_logq_kernel<32>) at the
    full 1e7 x 1024 size: oracle on a row subset, the CUDA-core kernel on every row;
  * the same path on CLUSTERED proposals (after 5 DM-PMC adaptations the 1024 proposals have collapsed onto a few
    samples: thousands of pairs sit inside the exact-refinement window of csrc/tc.cu) -- accuracy and time;
  * BASELINE configs[2] shape (8-D 16-mode target, 256 proposals) against fixtures produced by the reference's own
    M-PMC and LAIS adapt loops (tests/golden/make_golden.py::make_config3).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import ais_oracle as O

pytestmark = pytest.mark.gpu

LOGP_TOL = 1e-5
W_TOL = 1e-4


def gmm(means, sigmas, weights, support=None):
    from ais_benchmarks_b200.distributions import CGaussianMixtureModel
    dims = np.asarray(means).shape[1]
    if support is None:
        support = [np.full(dims, -1.0), np.full(dims, 1.0)]
    return CGaussianMixtureModel({"means": means, "sigmas": sigmas, "weights": weights, "support": support})


def _timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def test_full_size_config5_hinted_tensor_core_path():
    """The call bench.py times for configs[4]: _device.dm_weights(xs, target, None, proposals, hint=own) at 1e7 samples x
    1024 proposals, D = 32. ~264 row tiles per persistent CTA: the A-tile double buffer and the mbarrier parities of the
    tensor-core kernel run through hundreds of iterations."""
    import bench
    from ais_benchmarks_b200 import _device
    mu, sig, w, sup, pmeans, per = bench.dm_workload(scale=1.0)
    target = gmm(mu, sig, w, sup)
    pm_d = torch.from_numpy(pmeans).cuda().float().contiguous()
    qmix = _device.DeviceMixture.kde(pm_d, None, len(pmeans), 0.05)
    assert _device.tc_operand(qmix) is not None, "configs[4] must be eligible for the tensor-core path"
    _device.seed(1234)
    xs = _device.sample_iso(pm_d, per, 0.05)
    n = xs.shape[0]
    assert n > 9.9e6
    own = (torch.arange(n, device="cuda") // per).to(torch.int32)
    logw, logp, logq, part = _device.dm_weights(xs, target.device_mixture(), None, qmix, want_logp=True, want_logq=True,
                                                hint=own)
    p = _device.check_partials(part)[0]                       # raises if a tensor-core CTA aborted
    assert p[3] == n
    rows = np.random.RandomState(2).choice(n, size=256, replace=False)
    xr = xs[rows].double().cpu().numpy()
    pm32 = pm_d.double().cpu().numpy()
    assert rel_err(logq[rows].cpu().numpy(), O.iso_mixture_log_prob(xr, pm32, 0.05)) < LOGP_TOL, rel_err.last
    assert rel_err(logp[rows].cpu().numpy(), O.gmm_log_prob(xr, mu, sig, w)) < LOGP_TOL, rel_err.last
    # every one of the 1e7 rows against the CUDA-core kernel
    logq_c = _device.mixture_logpdf(xs, qmix)
    err = ((logq.double() - logq_c.double()).abs() / logq_c.double().abs().clamp(min=1.0)).max()
    assert float(err) <= 2e-6, float(err)
    lw = logw.double()
    M = float(lw.max())
    assert p[0] == pytest.approx(M, abs=1e-6) and p[1] == pytest.approx(float(torch.exp(lw - M).sum()), rel=1e-5)


@pytest.mark.parametrize("D,modes", [(32, 64), (16, 16)])
def test_tensor_core_path_on_clustered_proposals(D, modes, capsys):
    """Worst case for the screen: after adaptation DM-PMC proposals are (near-)duplicates of a few heavy samples, so for
    every sample hundreds of proposals lie within the 2^-30 refine window and are re-evaluated exactly. The result must
    still match the CUDA-core kernel to rounding and the oracle to 1e-5; the time of both kernels is printed."""
    import bench
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CDeterministicMixtureAIS
    N, K = 1024, 40
    mu, sig, w, sup = bench.random_gmm(10 + D, D, modes)
    target = gmm(mu, sig, w, sup)
    np.random.seed(5)
    _device.seed(5)
    s = CDeterministicMixtureAIS({"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000,
                                  "sigma": 0.05, "cuda_graph": False, "device_outputs": True})
    x, wn = s.importance_sample(target, 5 * N * K)            # 5 adapt iterations through the hinted path
    assert x.shape[0] == 5 * N * K and abs(float(wn.double().sum()) - 1) < 1e-4
    pm = s.proposal_means().clone()
    n_unique = int(torch.unique(pm, dim=0).shape[0])
    qmix = _device.DeviceMixture.kde(pm, None, N, 0.05)
    op = _device.TensorCoreOperand(qmix)                       # force the tensor-core path ...
    assert op.crowding > _device.TC_MAX_CROWDING               # ... which the dispatcher no longer takes for such proposals
    assert _device.tc_operand(qmix) is None
    per = 300                                                  # 307 200 samples: ~8 row tiles per persistent CTA
    xs = _device.sample_iso(pm, per, 0.05)
    hint = (torch.arange(xs.shape[0], device="cuda") // per).to(torch.int32)
    got = _device.tc_logq(xs, op, hint)
    ref32 = _device.mixture_logpdf(xs, qmix)
    # hundreds of comparable terms per sample: the two kernels add them in different orders (four float32 partial sums
    # here, float32 chunks folded into float64 there); measured: each within 5.5e-6 of the float64 oracle, 6.6e-6 apart
    assert rel_err(got.cpu().numpy(), ref32.double().cpu().numpy()) < 1e-5, rel_err.last
    rows = np.random.RandomState(D).choice(xs.shape[0], size=512, replace=False)
    ref64 = O.iso_mixture_log_prob(xs[rows].double().cpu().numpy(), pm.double().cpu().numpy(), 0.05)
    assert rel_err(got[rows].cpu().numpy(), ref64) < LOGP_TOL, rel_err.last
    out = torch.empty_like(got)
    t_tc = _timed(lambda: _device.tc_logq(xs, op, hint, out))
    t_cc = _timed(lambda: _device.mixture_logpdf(xs, qmix, out))
    # the same number of samples from UN-adapted (uniform) proposals, for scale
    pm0 = torch.from_numpy(np.random.RandomState(1).uniform(sup[0], sup[1], size=(N, D))).cuda().float().contiguous()
    q0 = _device.DeviceMixture.kde(pm0, None, N, 0.05)
    x0 = _device.sample_iso(pm0, per, 0.05)
    assert _device.tc_operand(q0) is not None and _device.tc_operand(q0).crowding < _device.TC_MAX_CROWDING
    t_tc0 = _timed(lambda: _device.tc_logq(x0, _device.tc_operand(q0), hint, out))
    with capsys.disabled():
        print("\n[clustered proposals D=%d] %d distinct means of %d; tc_logq %.3f ms (uniform proposals: %.3f ms), "
              "CUDA-core kernel %.3f ms, %d samples" % (D, n_unique, N, t_tc, t_tc0, t_cc, xs.shape[0]))


@pytest.mark.parametrize("update", ["host", "device"])
@pytest.mark.parametrize("geom", ["std", "tiny"])
def test_mpmc_config3_shape_against_reference_iterations(geom, update):
    """Two consecutive iterations of the reference's CMixturePMC at the BASELINE configs[2] shape (D = 8, 16-mode target,
    N = 256 proposals; make_config3): batch-normalised weights, alpha, mu and Sigma (incl. which proposals hit the 10*I
    failsafe of quirk Q3). "tiny" iteration 0 produces genuine full covariances for all 256 proposals; its iteration 1
    then samples from / weighs against those FULL-covariance proposals. update = "host": the float64 numpy update of
    CMixturePMC.adapt; "device": aisb_mpmc_update_f64 (CMixturePMC.adapt_device), the path importance_sample takes."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CMixturePMC
    g = load_golden("config3")
    half = float(g["half_" + geom])
    sup = [np.full(8, -half), np.full(8, half)]
    target = gmm(g["tmeans_" + geom], g["tsigmas_" + geom], g["tweights_" + geom], sup)
    saw_full, saw_failsafe = False, False
    for it in range(2):
        t = "_%s_%d" % (geom, it)
        x = g["x" + t]
        K, D = x.shape
        N = len(g["means" + t])
        assert (D, N) == (8, 256) and K >= 4000
        np.random.seed(0)
        s = CMixturePMC({"space_min": sup[0], "space_max": sup[1], "dims": D, "K": K, "N": N, "J": 1000,
                         "sigma": float(g["sigma_" + geom])})
        s._means, s._covs = g["means" + t].copy(), g["covs" + t].copy()
        s.wproposals = g["alpha" + t].copy()
        s._mix_weights = g["mixw" + t].copy()
        s._mixture = None
        xd = torch.from_numpy(x).cuda().float()
        logw, logq, part = s.weigh(xd, target)
        ph = part.cpu().numpy()
        # log densities behind the weights, against CMixtureModel.log_prob of the reference on the first 1000 draws
        # (DIAG proposals in iteration 0, FULL-covariance / failsafe 10*I proposals in iteration 1)
        nr = len(g["lq" + t])
        assert rel_err(logq[:nr].cpu().numpy(), g["lq" + t]) < LOGP_TOL, rel_err.last
        assert rel_err((logw + logq)[:nr].cpu().numpy(), g["lp" + t]) < LOGP_TOL, rel_err.last
        ref_w = g["w" + t]
        if not np.isfinite(ref_w).any():
            # "tiny" iteration 1: the proposals the reference adapted to are ~3e4 times wider than the target, every
            # linear-domain weight exp(pi - q) underflows to 0 and w / sum(w) is NaN for the whole batch
            # (m_pmc.py:96,108). The log-domain weights here stay finite; nothing further to compare.
            assert np.isfinite(logw.cpu().numpy()).all()
            continue
        w = torch.exp(logw.double() - _device.weight_logsum(ph)).cpu().numpy()
        big = ref_w > 1e-9 * ref_w.max()
        assert big.sum() > 0 and np.max(np.abs(w[big] - ref_w[big]) / ref_w[big]) < W_TOL
        assert abs(w.sum() - ref_w.sum()) < 1e-6
        if update == "host":
            s.adapt(xd, logw, logq, ph)
        else:
            s.adapt_device(xd, logw, logq, part, K)
            assert s._dev_valid and s._host is None            # parameters live in HBM until somebody reads them
            assert s.proposal_mixture().kind == (1 if g["failsafe" + t].all() else 2)   # DIAG after the failsafe, else FULL
        assert np.allclose(s.wproposals, g["new_alpha" + t], rtol=W_TOL, atol=1e-9)
        live = g["new_alpha" + t] > 1e-9
        assert live.sum() > 0
        assert np.allclose(s._means[live], g["new_means" + t][live], rtol=W_TOL, atol=1e-5 * half)
        fs = g["failsafe" + t]
        ours_fs = np.all(s._covs == np.diag(np.ones(D) * 10)[None], axis=(1, 2))
        assert np.array_equal(ours_fs[live], fs[live])
        ok = live & ~fs
        if ok.any():
            scale = np.abs(g["new_covs" + t][ok]).max()
            assert np.allclose(s._covs[ok], g["new_covs" + t][ok], rtol=1e-3, atol=1e-4 * scale)
        saw_full |= bool(ok.any())
        saw_failsafe |= bool(fs[live].any())
    assert saw_failsafe == (geom == "std")
    assert saw_full == (geom == "tiny")


def test_lais_config3_shape_against_reference_iteration():
    """One iteration of the reference's CLayeredAIS at the configs[2] shape (256 proposals x 39 draws, D = 8, L = 10 MH
    steps; make_config3): normalised DM weights and NESS on the reference's own samples, chain positions after the MH
    layer with the reference's captured noise."""
    from ais_benchmarks_b200 import _device
    from ais_benchmarks_b200.sampling_methods import CLayeredAIS
    g = load_golden("config3")
    K, L, sigma, mh_sigma = g["l_cfg"]
    K, L = int(K), int(L)
    target = gmm(g["l_tmeans"], g["l_tsigmas"], g["l_tweights"], [g["l_lo"], g["l_hi"]])
    before = g["l_before"]
    N, D = before.shape
    assert (N, D) == (256, 8)
    np.random.seed(1)
    s = CLayeredAIS({"space_min": g["l_lo"], "space_max": g["l_hi"], "dims": D, "K": K, "N": N, "J": 1, "L": L,
                     "sigma": float(sigma), "mh_sigma": float(mh_sigma)})
    s._set_means_host(before)
    xd = torch.from_numpy(g["l_x"]).cuda().float()
    logw, part = s.weigh(xd, target, hint=s._own_proposal(K))
    ph = part.cpu().numpy()
    w = torch.exp(logw.double() - _device.weight_logsum(ph)).cpu().numpy()
    ref_w = g["l_w"]
    big = ref_w > 1e-9 * ref_w.max()
    assert np.max(np.abs(w[big] - ref_w[big]) / ref_w[big]) < W_TOL
    assert _device.weight_ess(ph) / len(ref_w) == pytest.approx(float(g["l_ness"]), rel=W_TOL)
    s.mcmc_step(target, torch.from_numpy(g["l_delta"]).cuda().float(), torch.from_numpy(g["l_u"]).cuda().float())
    after = s.proposal_means().cpu().numpy()
    assert np.allclose(after, g["l_after"], rtol=0, atol=2e-6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_staged_upload_is_numpy_astype_float32(dtype, monkeypatch):
    """aisb_upload_as_f32 (the host -> HBM leg of every numpy call: reference arrays are float64, README.md:41-43) is
    bit-identical to numpy's astype(float32) for every thread count, chunk boundary and special value, keeps the
    ordering contract with the caller's stream, and leaves the source untouched."""
    from ais_benchmarks_b200 import _device
    rs = np.random.RandomState(11)
    chunk = 1 << 20
    for count in (_device.HOST_CAST_MAX + 1, chunk - 1, chunk, chunk + 1, 3 * chunk + 12345, 9 * chunk + 7):
        x = (rs.standard_normal(count) * 10.0 ** rs.randint(-30, 30, count)).astype(dtype)
        x[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4028235e38, 1e300 if dtype == np.float64 else 1e30]
        keep = x.copy()
        with np.errstate(over="ignore"):
            want = x.astype(np.float32)
        for threads in (1, 2, 3, 5, 8):
            monkeypatch.setenv("AISB_UPLOAD_THREADS", str(threads))
            got = _device.upload_f32(x).cpu().numpy()
            assert got.dtype == np.float32 and got.shape == x.shape
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (count, threads)
        assert np.array_equal(x.view(np.uint8), keep.view(np.uint8))
    # 2-D points through the public entry point, queued behind work on the caller's stream that uses the same memory
    pts = rs.rand(300_000, 8)
    scratch = torch.empty(300_000, 8, device="cuda")
    for _ in range(3):
        scratch.normal_()                      # pending kernel on the current stream; its block may be handed out again
        del scratch
        dev, was_numpy = _device.to_device_points(pts, 8)
        assert was_numpy and torch.equal(dev.cpu(), torch.from_numpy(pts.astype(np.float32)))
        scratch = torch.empty(300_000, 8, device="cuda")
    with pytest.raises(_device.AisbError):
        _device.check(_device._lib.load().aisb_upload_as_f32(None, 1, 10, _device._ptr(scratch), 2, None))


def test_fast_kde_exact_reevaluation_plans_itself(monkeypatch):
    """aisb_mixture_logpdf_fast_f32 lists the points whose error estimate is too large and re-evaluates them with ONE
    fixed-size launch that plans itself from the list length on the device (csrc/mixture.cu: subset_plan). Every plan
    is exercised against the exact kernel on the same points: nothing or little listed; every point listed with fewer
    point tiles than CTAs (split component range + merge) and with MORE tiles than CTAs (the CTAs stride over the tiles
    and write directly). `AISB_FAST_DEBUG=2` (no recentring) on data far from the origin lists every point."""
    from ais_benchmarks_b200 import _device, _lib
    from ais_benchmarks_b200._device import _ptr, stream_ptr, check
    lib = _lib.load()
    monkeypatch.setattr(_device, "MORTON_MIN_ROWS", 1 << 16)
    rs = np.random.RandomState(77)
    D, K, bw = 8, (1 << 16) + 129, 0.02
    modes = rs.uniform(0.2, 0.8, size=(5, D))

    def draw(m, spread):
        return (modes[rs.randint(0, 5, size=m)] + spread * rs.standard_normal((m, D))).astype(np.float32)

    def run(n, offset, debug):
        c = draw(K, 0.1) + np.float32(offset)
        x = draw(n, 0.2) + np.float32(offset)
        cd, xd = torch.from_numpy(c).cuda(), torch.from_numpy(x).cuda()
        mix = _device.DeviceMixture.kde(cd, None, K, bw)
        assert mix.spatially_ordered
        xs = xd[_device.morton_order(xd)].contiguous()
        ws, sc = _device.workspace_for(n, K, D), _device._kde_fast_scratch(n)
        fast, exact = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
        stats = torch.zeros(3, dtype=torch.int64, device="cuda")
        if debug:
            monkeypatch.setenv("AISB_FAST_DEBUG", str(debug))
        else:
            monkeypatch.delenv("AISB_FAST_DEBUG", raising=False)
        check(lib.aisb_mixture_logpdf_fast_f32(_ptr(xs), n, mix.ref(), _ptr(fast), _ptr(ws), ws.numel(), _ptr(sc),
                                               sc.numel(), _ptr(stats), stream_ptr()))
        monkeypatch.delenv("AISB_FAST_DEBUG", raising=False)
        check(lib.aisb_mixture_logpdf_f32(_ptr(xs), n, mix.ref(), _ptr(exact), _ptr(ws), ws.numel(), stream_ptr()))
        torch.cuda.synchronize()
        a, b = fast.double().cpu().numpy(), exact.double().cpu().numpy()
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        dev = float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(1.0, np.abs(b[fin])))) if fin.any() else 0.0
        return int(stats[2]), dev

    tile = 512                                   # points per CTA of the re-evaluation kernel at D = 8
    listed, dev = run(70_000, 0.0, 0)            # recentred: a small minority listed (sparse warps: 5 % here)
    assert 0 < listed < 70_000 // 5 and dev < 5e-6, (listed, dev)
    listed, dev = run(70_000, 100.0, 2)          # all listed, 137 tiles < 4096 CTAs: 29-way split + merge
    assert listed == 70_000 and dev < 2e-6, (listed, dev)
    listed, dev = run(tile + 1, 100.0, 2)        # two tiles, the second with one point: 1024-way split (the cap)
    assert listed == tile + 1 and dev < 2e-6, (listed, dev)
    n_big = 4096 * tile + 3 * tile + 17          # more tiles than CTAs: unsplit, strided
    listed, dev = run(n_big, 100.0, 2)
    assert listed == n_big and dev < 2e-6, (listed, dev)


def test_gather_rows_kernel():
    """aisb_gather_rows_f32 (Z-ordering of point sets, block-cyclic shards) equals torch indexing for vector and scalar
    row widths, repeated indices and empty selections; out-of-range indices are skipped and counted."""
    from ais_benchmarks_b200 import _device, _lib
    from ais_benchmarks_b200._device import _ptr, stream_ptr, check
    g = torch.Generator(device="cuda").manual_seed(3)
    for D in (1, 3, 4, 8, 12, 32):
        x = torch.randn(100_003, D, device="cuda", generator=g)
        for m in (0, 1, 77_777, 250_000):
            idx = torch.randint(0, x.shape[0], (m,), device="cuda", generator=g)
            assert torch.equal(_device.gather_rows(x, idx), x[idx])
        view = x[1:]                                   # base no longer 16-byte aligned for D = 1, 3
        idx = torch.randperm(view.shape[0], device="cuda", generator=g)
        assert torch.equal(_device.gather_rows(view, idx), view[idx])
    x = torch.randn(1000, 8, device="cuda", generator=g)
    idx = torch.tensor([0, 999, 1000, -1, 5], device="cuda")
    out = torch.full((5, 8), 7.0, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    check(_lib.load().aisb_gather_rows_f32(_ptr(x), 1000, 8, _ptr(idx), 5, _ptr(out), _ptr(err), stream_ptr()))
    assert int(err) == 4                               # two bad rows x two float4 pieces each
    assert torch.equal(out[[0, 1, 4]], x[[0, 999, 5]]) and bool((out[2:4] == 7.0).all())


def test_staged_download_is_exact_widening(monkeypatch):
    """aisb_download_f32 (HBM float32 -> the float64 / float32 numpy arrays the reference's API returns) is the exact
    widening of the device values for every thread count and chunk boundary, waits for the producing kernel on the
    caller's stream, and does not write outside the destination."""
    from ais_benchmarks_b200 import _device
    g = torch.Generator(device="cuda").manual_seed(21)
    chunk = 1 << 20
    for count in (_device.HOST_CAST_MAX + 1, chunk - 1, chunk + 1, 3 * chunk + 12345, 9 * chunk + 7):
        t = torch.randn(count, device="cuda", generator=g) * 1e3
        t[:5] = torch.tensor([0.0, -0.0, float("inf"), float("nan"), 1e-45], device="cuda")
        want = t.cpu().numpy()
        for threads in (1, 2, 3, 5, 8):
            monkeypatch.setenv("AISB_UPLOAD_THREADS", str(threads))
            for dt in (np.float64, np.float32):
                got = _device.download(t, dt)
                assert got.dtype == dt and got.shape == (count,)
                assert np.array_equal(got.view(np.uint64 if dt == np.float64 else np.uint32),
                                      want.astype(dt).view(np.uint64 if dt == np.float64 else np.uint32)), (count, threads)
    # ordering with the producing kernel + 2-D shape + guard bytes around a caller-owned destination
    lib = _device._lib.load()
    src = torch.empty(700_000, 8, device="cuda")
    for rep in range(3):
        src.normal_(generator=g)                                     # still running when the download is queued
        got = _device.from_device(src, True)
        assert got.shape == (700_000, 8) and np.array_equal(got, src.cpu().numpy().astype(np.float64))
    buf = np.full(src.numel() + 16, 7.0)
    _device.check(lib.aisb_download_f32(_device._ptr(src), src.numel(), buf[8:].ctypes.data, 1, 4, _device.stream_ptr()))
    assert (buf[:8] == 7.0).all() and (buf[-8:] == 7.0).all()
    assert np.array_equal(buf[8:-8], src.cpu().numpy().astype(np.float64).reshape(-1))
    with pytest.raises(_device.AisbError):
        _device.check(lib.aisb_download_f32(_device._ptr(src), 10, None, 1, 2, None))
