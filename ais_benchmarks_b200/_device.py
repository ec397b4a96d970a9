"""Device plumbing between the Python API and libais_b200.so.

PyTorch is used for device memory, streams and (in parallel.py) torch.distributed only; every
arithmetic step of the hot path is a call into the C ABI (include/ais_b200.h). Nothing in here
falls back to the CPU: without a CUDA device or without the built library the calls raise.
"""
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import AisbError, PackedMixture, COV_ISO_SHARED, COV_DIAG, COV_FULL, check

LOG2E = 1.4426950408889634074
LOG_2PI = 1.8378770664093454836

_workspaces = {}


def require_cuda():
    if not torch.cuda.is_available():
        raise AisbError("ais_benchmarks_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return _lib.load()


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device():
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def check_points(x, dims):
    """Shape contract of CDistribution._check_shape (reference distributions/distributions.py:128-147):
    1-D input -> (1, dims); 2-D input must have shape[1] == dims; anything else raises ValueError."""
    shape = tuple(x.shape)
    if len(shape) == 1:
        if shape[0] != dims:
            raise ValueError("cannot reshape array of size %d into shape (1,%d)" % (shape[0], dims))
        return x.reshape(1, dims)
    if len(shape) == 2:
        if shape[1] != dims:
            raise ValueError("Shape of x does not match self.dims = %d. is Nx%d array that contains N points of %d "
                             "dimensions. x.shape=%s" % (dims, dims, dims, str(shape)))
        return x
    raise ValueError("Shape of x does not match self.dims = %d. is Nx%d array that contains N points of %d "
                     "dimensions. x.shape=%s" % (dims, dims, dims, str(shape)))


def to_device_points(x, dims):
    """-> (float32 contiguous CUDA tensor [n, dims], was_numpy). Accepts numpy arrays, lists or torch tensors."""
    if not is_tensor(x):
        x = np.asarray(x)
        if x.dtype == object:
            raise ValueError("Unsupported samples data format")
    x = check_points(x, dims)
    require_cuda()
    if is_tensor(x):
        return x.to(device=device(), dtype=torch.float32).contiguous(), False
    return upload_f32(x), True


HOST_CAST_MAX = 1 << 16   # elements up to which a float64 -> float32 cast on the host is cheaper than moving 8-byte values


def _upload_threads():
    """Host threads of the staged uploader: AISB_UPLOAD_THREADS, else the cores this rank can claim, at most 8. Sources
    larger than the host's last-level cache stream from DRAM: 64 MB of float64 take 2.6-4.0 ms with 4 threads and
    2.0-2.3 ms with 8 on the 16-core host of the B200 box (scripts/upload_timing.py)."""
    env = os.environ.get("AISB_UPLOAD_THREADS")
    if env:
        return max(1, min(8, int(env)))
    local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(8, (os.cpu_count() or 1) // local))


def gather_rows(x, idx):
    """x[idx] for a float32 CUDA matrix and an int64 CUDA index vector (aisb_gather_rows_f32). The caller vouches for
    the indices (permutations and shard positions computed here); torch's own indexing spends one block per row."""
    if x.dim() != 2 or x.dtype != torch.float32 or idx.dtype != torch.int64 or not x.is_cuda:
        return x[idx]
    x, idx = x.contiguous(), idx.contiguous()
    out = torch.empty((idx.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    check(_lib.load().aisb_gather_rows_f32(_ptr(x), x.shape[0], x.shape[1], _ptr(idx), idx.shape[0], _ptr(out), None,
                                           stream_ptr()), "aisb_gather_rows_f32")
    return out


def upload_f32(x, out=None):
    """numpy array -> float32 contiguous CUDA tensor (`out`: a contiguous float32 CUDA tensor of the same size to fill
    instead of a new one). The reference hands over float64 (README.md:41-43): small arrays are cast on the host and
    copied by torch; large float arrays go through the library's staged uploader (aisb_upload_as_f32: several host
    threads narrow chunks into pinned staging buffers while the previous chunks are in flight -- numpy's
    single-threaded astype runs at ~1.5 GB/s and a pageable copy at 7-15 GB/s, either would dominate a 1e6 x 8
    evaluation)."""
    x = np.ascontiguousarray(x)
    if out is not None:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == x.size
    if x.size <= HOST_CAST_MAX or x.dtype not in (np.float64, np.float32):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device(), non_blocking=False)
        if out is None:
            return t
        out.copy_(t.reshape(out.shape))
        return out
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=device())
    check(_lib.load().aisb_upload_as_f32(x.ctypes.data, 1 if x.dtype == np.float64 else 0, x.size, _ptr(out),
                                         _upload_threads(), stream_ptr()))
    return out


def download(t, dtype=np.float64):
    """float32 CUDA tensor -> fresh numpy array of `dtype` (float64: the reference's dtype; or float32). Large tensors
    go through the library's staged downloader (aisb_download_f32: host threads pull chunks through pinned buffers and
    widen them into the destination, first-touching its pages in parallel); small ones through torch."""
    t = t.detach()
    if t.numel() <= HOST_CAST_MAX or t.dtype != torch.float32 or not t.is_cuda:
        return t.cpu().numpy().astype(dtype)
    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=dtype)
    check(_lib.load().aisb_download_f32(_ptr(t), t.numel(), out.ctypes.data, 1 if dtype == np.float64 else 0,
                                        _upload_threads(), stream_ptr()), "aisb_download_f32")
    return out


def from_device(t, as_numpy):
    """float32 device vector -> numpy float64 (reference dtype) or the tensor itself."""
    return download(t) if as_numpy else t


def exp_from_device(t, as_numpy):
    """exp of a float32 device log-density vector: float64 numpy (so that values below the float32 range survive,
    as in the reference's float64 prob) or a float32 tensor."""
    if as_numpy:
        return torch.exp(t.detach().double()).cpu().numpy()
    return torch.exp(t)


def workspace(nbytes):
    dev = torch.cuda.current_device()
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device())
        _workspaces[dev] = ws
    return ws


def workspace_bytes_for(n, K, D):
    lib = require_cuda()
    need = C.c_size_t(0)
    check(lib.aisb_workspace_bytes(int(n), int(K), int(D), C.byref(need)), "aisb_workspace_bytes")
    return int(need.value)


def workspace_for(n, K, D):
    lib = require_cuda()
    need = C.c_size_t(0)
    check(lib.aisb_workspace_bytes(int(n), int(K), int(D), C.byref(need)), "aisb_workspace_bytes")
    return workspace(need.value)


# ---------------------------------------------------------------------------------------------------------------------
# Spatial (Morton / Z-order) ordering of large point sets. The pairwise kernels skip the log-sum-exp fold of a group of
# components when it is negligible for EVERY point of the warp (csrc/pairwise.cuh, lse_update2); that only fires when a
# warp's points and a group's components are spatially coherent, so a large KDE packs its centres in Morton order and
# evaluates its points in Morton order (results are scattered back: callers never see the permutation).
# ---------------------------------------------------------------------------------------------------------------------
MORTON_MIN_ROWS = 1 << 18     # centres: below this the sort costs more than the skipped folds can save
MORTON_MIN_PAIRS = 5e10       # points x centres from which the evaluation is worth re-ordering the points (~0.5 ms)
MORTON_MAX_DIMS = 16
_morton_luts = {}


def _morton_lut(D, bits, dev):
    key = (D, bits, dev)
    if key not in _morton_luts:
        v = np.arange(1 << bits, dtype=np.int64)
        lut = np.zeros((D, 1 << bits), dtype=np.int64)
        for d in range(D):
            for b in range(bits):
                lut[d] |= ((v >> b) & 1) << (b * D + d)
        _morton_luts[key] = torch.from_numpy(lut).to(dev)
    return _morton_luts[key]


def morton_bits(D):
    return max(1, min(8, 32 // D))


def morton_order_torch(x):
    """Reference implementation of morton_order with torch element-wise operations only (any device; the test oracle
    of the key kernel, and the path taken with AISB_MORTON_TORCH=1)."""
    n, D = x.shape
    bits = morton_bits(D)
    lo = x.min(dim=0).values
    span = (x.max(dim=0).values - lo).clamp_min(1e-30)
    q = ((x - lo) / span * float(1 << bits)).clamp_(0, (1 << bits) - 1).to(torch.int64)
    lut = _morton_lut(D, bits, x.device)
    code = lut[0][q[:, 0]]
    for d in range(1, D):
        code = code + lut[d][q[:, d]]
    return torch.argsort(code)


def morton_keys(x):
    """Signed 32-bit Z-order keys of the rows of x (float32 CUDA [n, D]) over their bounding box (C ABI
    aisb_morton_keys_auto_f32: bounding box + keys in one call, no torch reductions on the serial path)."""
    lib = require_cuda()
    n, D = x.shape
    x = x.contiguous()
    keys = torch.empty(n, dtype=torch.int32, device=x.device)
    box = torch.empty(2 * D, dtype=torch.int32, device=x.device)      # scratch: encoded per-dimension min / max
    check(lib.aisb_morton_keys_auto_f32(_ptr(x), n, D, morton_bits(D), _ptr(box), _ptr(keys), stream_ptr()),
          "aisb_morton_keys_auto_f32")
    return keys


def morton_order(x):
    """Permutation (int64 CUDA [n]) that puts the rows of x (float32 CUDA [n, D]) in Z-order over their bounding box:
    morton_bits(D) most significant bits per coordinate, interleaved; key kernel + one 32-bit sort. No host sync."""
    if os.environ.get("AISB_MORTON_TORCH") == "1" or not x.is_cuda:
        return morton_order_torch(x)
    return torch.sort(morton_keys(x)).indices


class DeviceMixture:
    """A packed Gaussian mixture resident in HBM (layout: include/ais_b200.h "Packed mixture layout")."""

    def __init__(self, rows, offsets, K, D, kind, iso_scale=0.0, uniform_offset=0.0):
        self.rows, self.offsets = rows, offsets
        self.K, self.D, self.kind = int(K), int(D), int(kind)
        self.iso_scale, self.uniform_offset = float(iso_scale), float(uniform_offset)
        self._struct = PackedMixture(_ptr(rows), _ptr(offsets), self.K, self.D, self.kind, self.iso_scale,
                                     self.uniform_offset)

    def ref(self):
        return C.byref(self._struct)

    @staticmethod
    def pack_host(means, cov, weights, kind):
        """float64 host parameters -> (rows float32 [KP, RF], offsets float32 [KP], iso_scale). No CUDA needed."""
        lib = _lib.load()
        means = np.ascontiguousarray(means, dtype=np.float64)
        K, D = means.shape
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        expect = {COV_ISO_SHARED: 1, COV_DIAG: K * D, COV_FULL: K * D * D}[kind]
        if cov.size != expect:
            raise ValueError("covariance parameter has %d elements, expected %d" % (cov.size, expect))
        if w is not None and w.size != K:
            raise ValueError("len(models)=%d , len(weights)=%d" % (K, w.size))
        RF = lib.aisb_packed_row_floats(kind, D)
        if RF == 0:
            raise AisbError("unsupported mixture shape: D=%d cov_kind=%d (D <= 64; full covariance D <= 32)" % (D, kind))
        KP = lib.aisb_padded_components(K)
        rows = np.empty((KP, RF), dtype=np.float32)
        offs = np.empty(KP, dtype=np.float32)
        iso = C.c_float(0.0)
        rc = lib.aisb_pack_mixture_f64(means.ctypes.data, cov.ctypes.data, None if w is None else w.ctypes.data, K, D,
                                       kind, rows.ctypes.data, offs.ctypes.data, C.byref(iso))
        if rc != 0:
            msg = lib.aisb_last_error().decode()
            # the reference asserts det > 0 (CMultivariateNormal.py:48,56)
            raise AssertionError(msg)
        return rows, offs, iso.value

    @classmethod
    def from_params(cls, means, cov, weights, kind):
        rows, offs, iso = cls.pack_host(means, cov, weights, kind)
        require_cuda()
        K, D = np.asarray(means).shape
        dev = device()
        return cls(torch.from_numpy(rows).to(dev), torch.from_numpy(offs).to(dev), K, D, kind, iso, 0.0)

    @classmethod
    def kde(cls, samples, idx, n_c, bw, spatial_order=True):
        """KDE of sampling_methods/base.py:162-179 built on device: centres samples[idx], covariance bw*I, weights 1/n_c."""
        lib = require_cuda()
        n, D = samples.shape
        DP = lib.aisb_padded_dims(D)
        KP = lib.aisb_padded_components(n_c)
        ordered = n_c >= MORTON_MIN_ROWS and D <= MORTON_MAX_DIMS and spatial_order
        if ordered:
            # a mixture does not care about the order of its components: pack them along a Z-order curve
            chosen = samples[:n_c] if idx is None else gather_rows(samples, idx)
            perm = morton_order(chosen)
            idx = perm if idx is None else idx[perm]
        rows = torch.empty((KP, DP), dtype=torch.float32, device=samples.device)
        check(lib.aisb_kde_pack_f32(_ptr(samples), n, D, _ptr(idx), int(n_c), float(bw), _ptr(rows), stream_ptr()),
              "aisb_kde_pack_f32")
        uo = -math.log(n_c) - 0.5 * D * (LOG_2PI + math.log(bw))
        mix = cls(rows, None, n_c, D, COV_ISO_SHARED, lib.aisb_iso_scale(float(bw)), uo)
        mix.spatially_ordered = bool(ordered)
        return mix


def kde_mode():
    """Kernel family for a spatially ordered KDE evaluated on spatially ordered points (AISB_KDE_MODE):
    "fast" (default)  recentred expanded form + exact re-evaluation of the points whose error estimate is too large
    "screen"          expanded-form screen + exact centred refinement of every group that matters
    "ordered"         centred form for every pair, fold skipping only (the round-1 kernel)"""
    return os.environ.get("AISB_KDE_MODE", "fast")


def _kde_kernel_applies(mix, x):
    return (mix.D in (2, 4, 8, 16) and mix.kind == COV_ISO_SHARED and mix.offsets is None and x.data_ptr() % 16 == 0)


def kde_screen(mix):
    """The screen vector of a spatially ordered KDE (C ABI aisb_kde_screen_pack_f32), built once per mixture."""
    if getattr(mix, "_screen", None) is None:
        lib = require_cuda()
        mix._screen = torch.empty(lib.aisb_kde_screen_floats(mix.K), dtype=torch.float32, device=mix.rows.device)
        check(lib.aisb_kde_screen_pack_f32(mix.ref(), _ptr(mix._screen), stream_ptr()), "aisb_kde_screen_pack_f32")
    return mix._screen


_fast_scratch = {}


def _kde_fast_scratch(n):
    lib = require_cuda()
    dev = torch.cuda.current_device()
    need = int(lib.aisb_kde_fast_scratch_bytes(int(n)))
    buf = _fast_scratch.get(dev)
    if buf is None or buf.numel() < need:
        buf = torch.empty(need, dtype=torch.uint8, device=device())
        _fast_scratch[dev] = buf
    return buf


def _logpdf_ordered(lib, x, n, mix, out, ws, stats=None):
    """Both sides spatially ordered: the fast / screened kernels where they apply, else the fold-skipping kernel."""
    mode = kde_mode() if _kde_kernel_applies(mix, x) else "ordered"
    if mode == "fast":
        sc = _kde_fast_scratch(n)
        check(lib.aisb_mixture_logpdf_fast_f32(_ptr(x), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(), _ptr(sc),
                                               sc.numel(), _ptr(stats), stream_ptr()), "aisb_mixture_logpdf_fast_f32")
    elif mode == "screen":
        check(lib.aisb_mixture_logpdf_screened_f32(_ptr(x), n, mix.ref(), _ptr(kde_screen(mix)), _ptr(out), _ptr(ws),
                                                   ws.numel(), _ptr(stats), stream_ptr()),
              "aisb_mixture_logpdf_screened_f32")
    else:
        check(lib.aisb_mixture_logpdf_ordered_f32(_ptr(x), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(),
                                                  stream_ptr()), "aisb_mixture_logpdf_ordered_f32")


def mixture_logpdf(x, mix, out=None, points_ordered=False, stats=None):
    """x: float32 CUDA [n, D]; returns float32 CUDA [n] natural-log densities (C ABI: aisb_mixture_logpdf_f32).
    points_ordered: the caller has already put the rows of x in Z-order (parallel.spatial_shard): with a spatially
    ordered mixture the screen-and-refine / fold-skipping kernels are used directly, without the internal sort / scatter.
    stats: optional zeroed int64 CUDA [3] that receives the fast / screened kernel's unit counters."""
    lib = require_cuda()
    n = x.shape[0]
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=x.device)
    if n == 0:
        return out
    ws = workspace_for(n, mix.K, mix.D)
    if points_ordered and getattr(mix, "spatially_ordered", False):
        _logpdf_ordered(lib, x, n, mix, out, ws, stats)
        return out
    if getattr(mix, "spatially_ordered", False) and n >= (1 << 16) and float(n) * mix.K >= MORTON_MIN_PAIRS:
        # components are packed in Z-order: evaluate the points in Z-order too (warp-coherent fold skipping), scatter back
        perm = morton_order(x)
        xs = gather_rows(x, perm)
        tmp = torch.empty(n, dtype=torch.float32, device=x.device)
        _logpdf_ordered(lib, xs, n, mix, tmp, ws, stats)
        out[perm] = tmp
        return out
    check(lib.aisb_mixture_logpdf_f32(_ptr(x), n, mix.ref(), _ptr(out), _ptr(ws), ws.numel(), stream_ptr()),
          "aisb_mixture_logpdf_f32")
    return out


class TensorCoreOperand:
    """Pre-arranged B operand of the tcgen05 cross-term path for an ISO_SHARED mixture without offsets."""

    def __init__(self, mix):
        lib = require_cuda()
        nfl = lib.aisb_tc_image_floats(mix.K, mix.D)
        if nfl == 0 or mix.kind != COV_ISO_SHARED or mix.offsets is not None:
            raise AisbError("tensor-core path needs an ISO_SHARED mixture without offsets and 5 <= D <= 32")
        self.mix = mix
        self.image = torch.empty(nfl, dtype=torch.float32, device=mix.rows.device)
        check(lib.aisb_tc_pack_f32(mix.ref(), _ptr(self.image), stream_ptr()), "aisb_tc_pack_f32")
        self._screen_error = None
        self._crowding = None

    def _read_trailer(self):
        t = self.image[-8:-6].cpu().numpy()           # one small D2H read for both statistics
        vmax = float(t[0])
        self._screen_error = vmax * vmax / 512.0
        self._crowding = float(t[1]) / self.mix.K if t[1] > 0 else 1.0

    @property
    def screen_error(self):
        """Worst-case absolute error (log2 units) of the TF32 screen for a point as far from the origin as the
        farthest component: 2^-9 max_j ||a mu_j||^2 (the image trailer holds the norm). The kernel widens each
        point's refine window by twice its own bound, so a large value costs time, never accuracy."""
        if self._screen_error is None:
            self._read_trailer()
        return self._screen_error

    @property
    def crowding(self):
        """Mean number of components (itself included) inside a component's exact-refinement window (2^-30 of its own
        term). 1 for well separated proposals; after DM-PMC resampling many proposals are copies of the same few heavy
        samples and the refinement re-evaluates all of them for every sample: measured at D = 32, 1024 proposals of
        which 149 are distinct, the tensor-core kernel takes 6.0 ms where the CUDA-core kernel takes 0.79 ms
        (un-adapted proposals: 0.15 vs 0.79 ms; scripts/tc_clustered_check.py)."""
        if self._crowding is None:
            self._read_trailer()
        return self._crowding


def tc_logq(x, operand, hint, out=None):
    """log q(x) of operand.mix through the tensor-core path (C ABI: aisb_tc_logq_f32). hint: int32 CUDA [n], index of
    the component expected to dominate each point (DM-AIS: the proposal the sample was drawn from)."""
    lib = require_cuda()
    n = x.shape[0]
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=x.device)
    err = torch.zeros(1, dtype=torch.int32, device=x.device)
    check(lib.aisb_tc_logq_f32(_ptr(x), n, operand.mix.ref(), _ptr(hint), _ptr(operand.image), _ptr(out), _ptr(err),
                               stream_ptr()), "aisb_tc_logq_f32")
    if int(err.item()) != 0:
        raise AisbError("tensor-core kernel aborted (%d CTAs): barrier protocol time-out" % int(err.item()))
    return out


TC_MAX_SCREEN_ERROR = 8.0  # beyond this most pairs would be re-evaluated exactly: the CUDA-core kernel is faster
TC_MAX_CROWDING = 1.5      # mean components per refinement window beyond which the CUDA-core kernel is faster


def tc_eligible(mix):
    """Tensor-core proposal path: equal-weight isotropic mixtures in 9..32 dimensions (below that the cross term is
    not a real contraction and the CUDA-core kernel wins)."""
    return mix.kind == COV_ISO_SHARED and mix.offsets is None and 9 <= mix.D <= 32


def tc_operand(mix):
    """The mixture's tensor-core operand (built once per parameter set), or None when the mixture is not eligible, sits so
    far from the origin that the TF32 screen would not discriminate (TensorCoreOperand.screen_error), or its components
    crowd each other's refinement windows (TensorCoreOperand.crowding: adapted / resampled proposals)."""
    if not tc_eligible(mix):
        return None
    if getattr(mix, "_tc_operand", None) is None:
        mix._tc_operand = TensorCoreOperand(mix)
    op = mix._tc_operand
    return op if op.screen_error <= TC_MAX_SCREEN_ERROR and op.crowding <= TC_MAX_CROWDING else None


def dm_weights(x, target, logp_in, proposals, want_logp=False, want_logq=False, hint=None, out_logw=None,
               out_partials=None, ws=None):
    """DM weights. Without `hint`: the fused CUDA-core kernel (C ABI aisb_dm_weights_f32). With `hint` (int32 CUDA [n],
    the proposal each sample was drawn from) and an eligible proposal mixture: proposal density on the tensor cores
    (aisb_dm_weights_hinted_f32). Returns (logw, logp|None, logq|None, partials[4] device f64). out_logw [n] float32 /
    out_partials [4] float64 (contiguous CUDA views) receive the results in place when given (streamed weighting). `ws`:
    a caller-owned workspace (uint8 CUDA tensor, see workspace_bytes_for) instead of the shared per-device one -- a
    captured CUDA graph must own every buffer it touches."""
    lib = require_cuda()
    operand = tc_operand(proposals) if hint is not None and x.shape[0] > 0 else None
    if operand is not None:
        return _dm_weights_hinted(lib, x, target, logp_in, proposals, operand, want_logp, want_logq, hint, out_logw,
                                  out_partials)
    n = x.shape[0]
    dev = x.device
    logw = torch.empty(n, dtype=torch.float32, device=dev) if out_logw is None else out_logw
    logp = torch.empty(n, dtype=torch.float32, device=dev) if want_logp else None
    logq = torch.empty(n, dtype=torch.float32, device=dev) if want_logq else None
    partials = torch.empty(4, dtype=torch.float64, device=dev) if out_partials is None else out_partials
    Kmax = max(proposals.K, target.K if target is not None else 1)
    if ws is None:
        ws = workspace_for(max(n, 1), Kmax, proposals.D)
    check(lib.aisb_dm_weights_f32(_ptr(x), n, target.ref() if target is not None else None, _ptr(logp_in),
                                  proposals.ref(), _ptr(logw), _ptr(logp), _ptr(logq), _ptr(partials), _ptr(ws),
                                  ws.numel(), stream_ptr()), "aisb_dm_weights_f32")
    return logw, logp, logq, partials


def _dm_weights_hinted(lib, x, target, logp_in, proposals, operand, want_logp, want_logq, hint, out_logw=None,
                       out_partials=None):
    n, dev = x.shape[0], x.device
    logw = torch.empty(n, dtype=torch.float32, device=dev) if out_logw is None else out_logw
    logp = torch.empty(n, dtype=torch.float32, device=dev) if want_logp else None
    logq = torch.empty(n, dtype=torch.float32, device=dev) if want_logq else None
    partials = torch.empty(4, dtype=torch.float64, device=dev) if out_partials is None else out_partials
    Kmax = max(proposals.K, target.K if target is not None else 1)
    ws = workspace_for(n, Kmax, proposals.D)
    hint = hint.to(device=dev, dtype=torch.int32).contiguous()
    check(lib.aisb_dm_weights_hinted_f32(_ptr(x), n, target.ref() if target is not None else None, _ptr(logp_in),
                                         proposals.ref(), _ptr(operand.image), _ptr(hint), _ptr(logw),
                                         _ptr(logp), _ptr(logq), _ptr(partials), _ptr(ws), ws.numel(), stream_ptr()),
          "aisb_dm_weights_hinted_f32")
    return logw, logp, logq, partials


def check_partials(p):
    """Raise if a kernel reported an abort through a weight record (n_valid = -1: a tensor-core CTA that timed out on
    its barrier protocol left log q unwritten; a peer-memory exchange that timed out poisons its record the same way).
    p: one record [4] or a stack of records [G, 4], host array or tensor (a tensor is read back here: call this where
    the host reads results anyway, it synchronises the stream)."""
    if is_tensor(p):
        p = p.detach().cpu().numpy()
    p = np.asarray(p, dtype=np.float64).reshape(-1, 4)
    if (p[:, 3] < 0).any():
        raise AisbError("weight record poisoned (n_valid < 0): a tensor-core kernel or a peer exchange aborted on a "
                        "barrier time-out; the log weights of that batch are not valid")
    return p


def mixture_component_draws(weights, n, dev):
    """n component indices of a mixture with host weights `weights` (CMixtureModel.py:79-83). When every weight has
    been zeroed (degenerate adaptation) the reference's np.random.multinomial(1, zeros) puts its one count in the
    LAST category, so every draw comes from the last component; replicated here."""
    w = np.asarray(weights, dtype=np.float64)
    if not (w.sum() > 0):
        return torch.full((int(n),), len(w) - 1, dtype=torch.int64, device=dev)
    return torch.multinomial(torch.from_numpy(w).to(dev), int(n), replacement=True)


def logw_partials(logw):
    lib = require_cuda()
    n = logw.shape[0]
    partials = torch.empty(4, dtype=torch.float64, device=logw.device)
    ws = workspace(32 * (n // 2048 + 2) + 512)
    check(lib.aisb_logw_partials_f32(_ptr(logw), n, _ptr(partials), _ptr(ws), ws.numel(), stream_ptr()),
          "aisb_logw_partials_f32")
    return partials


def normalize_weights(logw, partials_host):
    lib = require_cuda()
    n = logw.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=logw.device)
    p = (C.c_double * 4)(*[float(v) for v in partials_host])
    check(lib.aisb_normalize_weights_f32(_ptr(logw), n, p, _ptr(out), stream_ptr()), "aisb_normalize_weights_f32")
    return out


def merge_weight_records(records):
    """records: float64 CUDA [G, 4] (one weight record per rank / streamed chunk) -> merged record, float64 CUDA [4].
    Device-side counterpart of aisb_weight_partials_merge: no host synchronisation (C ABI aisb_weight_records_merge)."""
    lib = require_cuda()
    records = records.reshape(-1, 4).contiguous()
    out = torch.empty(4, dtype=torch.float64, device=records.device)
    check(lib.aisb_weight_records_merge(_ptr(records), records.shape[0], _ptr(out), stream_ptr()),
          "aisb_weight_records_merge")
    return out


def merge_kld_records(records):
    """records: float64 CUDA [G, 8] -> merged KLD record, float64 CUDA [8] (C ABI aisb_kld_records_merge)."""
    lib = require_cuda()
    records = records.reshape(-1, 8).contiguous()
    out = torch.empty(8, dtype=torch.float64, device=records.device)
    check(lib.aisb_kld_records_merge(_ptr(records), records.shape[0], _ptr(out), stream_ptr()),
          "aisb_kld_records_merge")
    return out


def normalize_weights_dev(logw, records, out=None):
    """Normalised weights from weight record(s) that stay in HBM (C ABI aisb_normalize_weights_dev_f32). `records`:
    float64 CUDA [4] or [G, 4] (one per rank / streamed chunk; folded inside the kernel).
    -> (w float32 CUDA [n], stats float64 CUDA [3] = {log sum w, ESS, n_finite}); nothing is read back here."""
    lib = require_cuda()
    n = logw.shape[0]
    records = records.reshape(-1, 4)
    if not records.is_contiguous():
        records = records.contiguous()
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=logw.device)
    stats = torch.empty(3, dtype=torch.float64, device=logw.device)
    check(lib.aisb_normalize_weights_dev_f32(_ptr(logw), n, _ptr(records), records.shape[0], _ptr(out), _ptr(stats),
                                             stream_ptr()), "aisb_normalize_weights_dev_f32")
    return out, stats


_copy_streams = {}


def _copy_stream():
    dev = torch.cuda.current_device()
    if dev not in _copy_streams:
        _copy_streams[dev] = torch.cuda.Stream(device=dev)
    return _copy_streams[dev]


def dm_weights_streamed(x_host, target, proposals, hint=None, chunk_rows=None, x_out=None):
    """DM weights of a HOST-resident (ideally pinned) float32 sample matrix [n, D]: the rows are copied to HBM in chunks
    on a copy stream while the weight kernel works on the previous chunk, so the PCIe transfer hides the arithmetic.
    Every chunk leaves one weight record. -> (x_dev, logw, records [chunks, 4] CUDA f64; normalize_weights_dev and
    merge_weight_records fold them on the device).
    `hint`: int32 CUDA [n] as in dm_weights."""
    require_cuda()
    n, D = x_host.shape
    dev = device()
    xd = torch.empty((n, D), dtype=torch.float32, device=dev) if x_out is None else x_out
    logw = torch.empty(n, dtype=torch.float32, device=dev)
    if chunk_rows is None:
        # ~8 chunks, each a multiple of 1024 rows (keeps every chunk 16-byte aligned and tile-aligned), >= 64k rows
        chunk_rows = max(65536, ((n + 7) // 8 + 1023) // 1024 * 1024)
    bounds = list(range(0, n, chunk_rows)) + [n]
    nchunks = max(1, len(bounds) - 1)
    records = torch.empty((nchunks, 4), dtype=torch.float64, device=dev)
    main, side = torch.cuda.current_stream(), _copy_stream()
    side.wait_stream(main)           # xd / logw allocations and earlier users of a recycled block
    events = []
    with torch.cuda.stream(side):
        for a, b in zip(bounds[:-1], bounds[1:]):
            xd[a:b].copy_(x_host[a:b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
            events.append(ev)
    if n == 0:
        dm_weights(xd, target, None, proposals, hint=hint, out_logw=logw, out_partials=records[0])
    for c, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        main.wait_event(events[c])
        dm_weights(xd[a:b], target, None, proposals, hint=None if hint is None else hint[a:b], out_logw=logw[a:b],
                   out_partials=records[c])
    xd.record_stream(side)
    return xd, logw, records


def kld_partials(lp, lq):
    lib = require_cuda()
    n = lp.shape[0]
    partials = torch.empty(8, dtype=torch.float64, device=lp.device)
    ws = workspace(64 * (n // 2048 + 2) + 512)
    check(lib.aisb_kld_partials_f32(_ptr(lp), _ptr(lq), n, _ptr(partials), _ptr(ws), ws.numel(), stream_ptr()),
          "aisb_kld_partials_f32")
    return partials


def pair_partials(lp, lq, floor):
    """kld_partials with the inlier set {lp > floor and lq > floor} (C ABI aisb_pair_partials_f32)."""
    lib = require_cuda()
    n = lp.shape[0]
    partials = torch.empty(8, dtype=torch.float64, device=lp.device)
    ws = workspace(64 * (n // 2048 + 2) + 512)
    check(lib.aisb_pair_partials_f32(_ptr(lp), _ptr(lq), n, float(floor), _ptr(partials), _ptr(ws), ws.numel(),
                                     stream_ptr()), "aisb_pair_partials_f32")
    return partials


def pair_log_norms(partials_host):
    """(Lp, Lq) = inlier log-sum-exps of a (merged) pair-partials record; (-inf, -inf) for an empty inlier set."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p = partials_host
        if p[5] <= 0:
            return -np.inf, -np.inf
        return float(p[0] + np.log(p[1])), float(p[3] + np.log(p[4]))


def jsd_terms(lp, lq, floor, log_norm_p, log_norm_q):
    """Device tensor {sum of JSD terms in nats, dropped NaN terms} (C ABI aisb_jsd_terms_f32)."""
    lib = require_cuda()
    out = torch.empty(2, dtype=torch.float64, device=lp.device)
    check(lib.aisb_jsd_terms_f32(_ptr(lp), _ptr(lq), lp.shape[0], float(floor), float(log_norm_p), float(log_norm_q),
                                 _ptr(out), stream_ptr()), "aisb_jsd_terms_f32")
    return out


def weighted_mean(x, logw, log_norm):
    """sum_i x_i exp(logw_i - log_norm) as a float64 device tensor [D] (C ABI aisb_weighted_mean_f32)."""
    lib = require_cuda()
    n, D = x.shape
    out = torch.empty(D, dtype=torch.float64, device=x.device)
    check(lib.aisb_weighted_mean_f32(_ptr(x), n, D, _ptr(logw), float(log_norm), _ptr(out), stream_ptr()),
          "aisb_weighted_mean_f32")
    return out


def kld_finalize(partials_host):
    lib = _lib.load()
    p = (C.c_double * 8)(*[float(v) for v in partials_host])
    return float(lib.aisb_kld_finalize(p))


def kld_merge(a, b):
    lib = _lib.load()
    pa = (C.c_double * 8)(*[float(v) for v in a])
    pb = (C.c_double * 8)(*[float(v) for v in b])
    out = (C.c_double * 8)()
    lib.aisb_kld_partials_merge(pa, pb, out)
    return np.array(list(out), dtype=np.float64)


def weight_merge(a, b):
    lib = _lib.load()
    pa = (C.c_double * 4)(*[float(v) for v in a])
    pb = (C.c_double * 4)(*[float(v) for v in b])
    out = (C.c_double * 4)()
    lib.aisb_weight_partials_merge(pa, pb, out)
    return np.array(list(out), dtype=np.float64)


def weight_ess(p):
    lib = _lib.load()
    return float(lib.aisb_weight_partials_ess((C.c_double * 4)(*[float(v) for v in p])))


def weight_logsum(p):
    lib = _lib.load()
    return float(lib.aisb_weight_partials_logsum((C.c_double * 4)(*[float(v) for v in p])))


def mpmc_moments(x, proposals, logq, logw, log_norm, second=False):
    """-> float64 CUDA [K, D+1]: column 0 = sum_i w~_i rho_di, columns 1.. = sum_i w~_i rho_di x_i. With second=True
    (D <= 8) the row continues with the lower triangle of sum_i w~_i rho_di x_i x_i^T: [K, 1 + D + D (D+1) / 2]."""
    lib = require_cuda()
    D = proposals.D
    width = D + 1 + (D * (D + 1) // 2 if second else 0)
    out = torch.zeros((proposals.K, width), dtype=torch.float64, device=x.device)
    fn = lib.aisb_mpmc_moments2_f32 if second else lib.aisb_mpmc_moments_f32
    check(fn(_ptr(x), x.shape[0], proposals.ref(), _ptr(logq), _ptr(logw), float(log_norm), _ptr(out), stream_ptr()),
          "aisb_mpmc_moments2_f32" if second else "aisb_mpmc_moments_f32")
    return out


def raw_moments(x):
    """-> (sum x [D], sum x x^T [D, D]) float64 CUDA."""
    lib = require_cuda()
    n, D = x.shape
    out = torch.zeros(D + D * D, dtype=torch.float64, device=x.device)
    check(lib.aisb_raw_moments_f32(_ptr(x), n, D, _ptr(out), stream_ptr()), "aisb_raw_moments_f32")
    return out[:D], out[D:].reshape(D, D)


def lais_mh(means, target, delta, u, lo, hi):
    """In-place: N chains x L Metropolis-Hastings steps on `means` [N, D] (C ABI: aisb_lais_mh_f32)."""
    lib = require_cuda()
    N, D = means.shape
    L = delta.shape[0]
    check(lib.aisb_lais_mh_f32(_ptr(means), N, target.ref(), _ptr(delta), _ptr(u), L, _ptr(lo), _ptr(hi),
                               stream_ptr()), "aisb_lais_mh_f32")
    return means


def box_mixture_logpdf(x, lo, hi, logw_over_vol, open_upper):
    lib = require_cuda()
    n, D = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    check(lib.aisb_box_mixture_logpdf_f32(_ptr(x), n, D, _ptr(lo), _ptr(hi), _ptr(logw_over_vol), lo.shape[0],
                                          1 if open_upper else 0, _ptr(out), stream_ptr()),
          "aisb_box_mixture_logpdf_f32")
    return out


_rng_state = {"seed": 0x5EED_B200, "offset": 0, "rank": None}
_RANK_MIX = 0xD1B54A32D192ED03  # odd 64-bit constant: rank r draws from the Philox key  seed ^ (r * _RANK_MIX)


def _rank():
    """Rank of this process under torch.distributed (0 when not initialised)."""
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _key():
    """Philox key of THIS process. Under torch.distributed every rank folds its rank into the user's seed, so that
    sharded samplers, which draw each rank's share with the same (seed, offset) bookkeeping, never produce identical
    rows on two ranks (that would count every sample world_size times in the merged weights, ESS and moments). Host
    draws that must agree across ranks (initial centres, shard counts) come from numpy on rank 0 + a broadcast."""
    _sync_rank()
    return (_rng_state["seed"] ^ ((_rng_state["rank"] * _RANK_MIX) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def _sync_rank():
    """(Re-)derive the per-rank streams when the rank becomes known or changes (seed() may run before
    init_process_group): torch's CUDA generator and the device-resident Philox state follow the rank-folded key."""
    r = _rank()
    if _rng_state["rank"] == r:
        return
    _rng_state["rank"] = r
    key = (_rng_state["seed"] ^ ((r * _RANK_MIX) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    if torch.cuda.is_available():
        # resampling / MH acceptance draw their uniforms from torch's CUDA generator: one seed reproduces a whole run
        torch.cuda.manual_seed(key & 0x7FFFFFFFFFFFFFFF)
        for dev, st in _rng_states.items():   # in place: captured graphs keep their pointer
            st.copy_(torch.tensor([0, _graph_key(key)], dtype=torch.int64), non_blocking=False)


def seed(s):
    """Seed the device sampling kernels (Philox, host-offset and device-resident streams) and torch's CUDA generator
    (resampling uniforms). Streams are not comparable with numpy's. Under torch.distributed the rank is folded into
    every stream (see _key): one seed still reproduces a whole multi-rank run."""
    _rng_state["seed"] = int(s) & 0xFFFFFFFFFFFFFFFF
    _rng_state["offset"] = 0
    _rng_state["rank"] = None                 # force _sync_rank to re-derive torch's generator and the device streams
    _sync_rank()


def _next_offset(n_values):
    off = _rng_state["offset"]
    _rng_state["offset"] = off + (n_values + 3) // 4 + 1
    return off


def sample_iso(means, per, var):
    """means [K, D] float32 CUDA -> [K*per, D]: `per` draws from N(means[k], var*I) for every k, k-major."""
    lib = require_cuda()
    K, D = means.shape
    out = torch.empty((K * per, D), dtype=torch.float32, device=means.device)
    off = _next_offset(K * per * D)
    check(lib.aisb_sample_iso_f32(_ptr(means), K, D, int(per), float(var), _key(), off, _ptr(out),
                                  stream_ptr()), "aisb_sample_iso_f32")
    return out


_GRAPH_STREAM_KEY = 0x9E3779B97F4A7C15  # key offset of the device-resident stream (disjoint from the host-offset stream)
_rng_states = {}


def _graph_key(key=None):
    k = ((_key() if key is None else key) ^ _GRAPH_STREAM_KEY) & 0xFFFFFFFFFFFFFFFF
    return k - (1 << 64) if k >= (1 << 63) else k       # as a two's-complement int64


def rng_state():
    """Device-resident Philox state {position, key} used inside captured CUDA graphs (one int64[2] per device). seed()
    rewrites it IN PLACE, so graphs captured earlier follow a re-seed."""
    dev = torch.cuda.current_device()
    _sync_rank()
    if dev not in _rng_states:
        _rng_states[dev] = torch.tensor([0, _graph_key()], dtype=torch.int64, device=device())
    return _rng_states[dev]


def sample_iso_ctr(means, per, var, out=None):
    """sample_iso with the Philox state in device memory (C ABI aisb_sample_iso_ctr_f32): safe to capture in a CUDA
    graph -- every replay draws fresh numbers."""
    lib = require_cuda()
    K, D = means.shape
    if out is None:
        out = torch.empty((K * per, D), dtype=torch.float32, device=means.device)
    check(lib.aisb_sample_iso_ctr_f32(_ptr(means), K, D, int(per), float(var), _ptr(rng_state()), _ptr(out),
                                      stream_ptr()), "aisb_sample_iso_ctr_f32")
    return out


def sample_mixture(means, chol, idx):
    """means [K, D], chol [K, D, D] (lower-triangular Cholesky factors) float32 CUDA, idx int64 CUDA [n] -> [n, D] draws
    means[idx] + chol[idx] z (C ABI aisb_sample_mixture_f32)."""
    lib = require_cuda()
    K, D = means.shape
    n = idx.shape[0]
    out = torch.empty((n, D), dtype=torch.float32, device=means.device)
    # the kernel consumes ceil(DP / 4) Philox counters per draw, DP = D padded to a power of two (csrc/misc.cu,
    # sample_mixture_kernel): advance the stream by exactly that, or consecutive calls would share normal deviates
    counters_per_draw = (int(lib.aisb_padded_dims(D)) + 3) // 4
    off = _next_offset(n * counters_per_draw * 4)
    check(lib.aisb_sample_mixture_f32(_ptr(means), _ptr(chol), K, D, _ptr(idx), n, _key(), off, _ptr(out),
                                      stream_ptr()), "aisb_sample_mixture_f32")
    return out


def uniform_box(lo, hi, n):
    """lo, hi [D] float32 CUDA -> [n, D] uniform draws."""
    lib = require_cuda()
    D = lo.shape[0]
    out = torch.empty((n, D), dtype=torch.float32, device=lo.device)
    off = _next_offset(n * D)
    check(lib.aisb_uniform_box_f32(_ptr(lo), _ptr(hi), D, int(n), _key(), off, _ptr(out), stream_ptr()),
          "aisb_uniform_box_f32")
    return out
