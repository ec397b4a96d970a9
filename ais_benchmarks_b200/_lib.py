"""ctypes binding of libais_b200.so (C ABI declared in include/ais_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C ais_benchmarks_b200/csrc``.
There is deliberately NO fallback: if the shared object is missing, every compute entry point of
this package raises (north_star: "no CPU fallback").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AISB_LIB selects an alternative build of the same ABI (kernel tuning experiments); the default is the in-tree library
LIB_PATH = os.environ.get("AISB_LIB") or os.path.join(_HERE, "libais_b200.so")

AISB_OK = 0
COV_ISO_SHARED, COV_DIAG, COV_FULL = 0, 1, 2
LOG_TINY = -745.13  # AISB_LOG_TINY: log-domain image of "prob > 0" in float64
COMP_ALIGN = 32
MAX_DIMS = 64


class PackedMixture(C.Structure):
    """struct aisb_packed_mixture (include/ais_b200.h)."""
    _fields_ = [("d_rows", C.c_void_p), ("d_offsets", C.c_void_p), ("K", C.c_int64), ("D", C.c_int32),
                ("cov_kind", C.c_int32), ("iso_scale", C.c_float), ("uniform_offset", C.c_float)]


class AisbError(RuntimeError):
    pass


_vp, _i64, _i32, _u64, _dbl, _sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_size_t
_pm = C.POINTER(PackedMixture)

# name -> (restype, argtypes); mirrors include/ais_b200.h one to one
PROTOTYPES = {
    "aisb_abi_version": (C.c_int, []),
    "aisb_last_error": (C.c_char_p, []),
    "aisb_padded_dims": (_i32, [_i32]),
    "aisb_packed_row_floats": (_i64, [_i32, _i32]),
    "aisb_padded_components": (_i64, [_i64]),
    "aisb_iso_scale": (C.c_float, [_dbl]),
    "aisb_pack_mixture_f64": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, C.POINTER(C.c_float)]),
    "aisb_kde_pack_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _dbl, _vp, _vp]),
    "aisb_workspace_bytes": (C.c_int, [_i64, _i64, _i32, C.POINTER(_sz)]),
    "aisb_mixture_logpdf_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _sz, _vp]),
    "aisb_mixture_logpdf_ordered_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _sz, _vp]),
    "aisb_kde_screen_floats": (_i64, [_i64]),
    "aisb_kde_screen_pack_f32": (C.c_int, [_pm, _vp, _vp]),
    "aisb_mixture_logpdf_screened_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _vp, _sz, _vp, _vp]),
    "aisb_kde_fast_scratch_bytes": (_sz, [_i64]),
    "aisb_mixture_logpdf_fast_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "aisb_dm_weights_f32": (C.c_int, [_vp, _i64, _pm, _vp, _pm, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aisb_logw_partials_f32": (C.c_int, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "aisb_normalize_weights_f32": (C.c_int, [_vp, _i64, C.POINTER(_dbl), _vp, _vp]),
    "aisb_weight_records_merge": (C.c_int, [_vp, _i64, _vp, _vp]),
    "aisb_kld_records_merge": (C.c_int, [_vp, _i64, _vp, _vp]),
    "aisb_weight_exchange_doubles": (_i64, [_i32]),
    "aisb_weight_records_exchange": (C.c_int, [_vp, _vp, _i32, _i32, _u64, _vp, _vp]),
    "aisb_normalize_weights_dev_f32": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "aisb_weight_partials_merge": (None, [C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    "aisb_weight_partials_ess": (_dbl, [C.POINTER(_dbl)]),
    "aisb_weight_partials_logsum": (_dbl, [C.POINTER(_dbl)]),
    "aisb_kld_partials_f32": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "aisb_kld_partials_merge": (None, [C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    "aisb_kld_finalize": (_dbl, [C.POINTER(_dbl)]),
    "aisb_pair_partials_f32": (C.c_int, [_vp, _vp, _i64, C.c_float, _vp, _vp, _sz, _vp]),
    "aisb_jsd_terms_f32": (C.c_int, [_vp, _vp, _i64, C.c_float, _dbl, _dbl, _vp, _vp]),
    "aisb_weighted_mean_f32": (C.c_int, [_vp, _i64, _i32, _vp, _dbl, _vp, _vp]),
    "aisb_mpmc_moments_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _dbl, _vp, _vp]),
    "aisb_mpmc_moments2_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _dbl, _vp, _vp]),
    "aisb_mpmc_moments_dev_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _vp, _vp, _vp]),
    "aisb_mpmc_update_f64": (C.c_int, [_vp, _vp, _dbl, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       _vp, _vp]),
    "aisb_raw_moments_f32": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "aisb_lais_mh_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _i32, _vp, _vp, _vp]),
    "aisb_box_mixture_logpdf_f32": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "aisb_morton_keys_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "aisb_morton_keys_auto_f32": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "aisb_gather_rows_f32": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp]),
    "aisb_download_f32": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _vp]),
    "aisb_upload_as_f32": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _vp]),
    "aisb_sample_mixture_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _i64, _u64, _u64, _vp, _vp]),
    "aisb_sample_iso_ctr_f32": (C.c_int, [_vp, _i64, _i32, _i64, _dbl, _vp, _vp, _vp]),
    "aisb_sample_iso_f32": (C.c_int, [_vp, _i64, _i32, _i64, _dbl, _u64, _u64, _vp, _vp]),
    "aisb_uniform_box_f32": (C.c_int, [_vp, _vp, _i32, _i64, _u64, _u64, _vp, _vp]),
    "aisb_tc_image_floats": (_i64, [_i64, _i32]),
    "aisb_tc_pack_f32": (C.c_int, [_pm, _vp, _vp]),
    "aisb_tc_logq_f32": (C.c_int, [_vp, _i64, _pm, _vp, _vp, _vp, _vp, _vp]),
    "aisb_dm_weights_hinted_f32": (C.c_int, [_vp, _i64, _pm, _vp, _pm, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "aisb_peak_fma_f32": (C.c_int, [_i64, _vp, C.POINTER(_dbl), _vp]),
    "aisb_peak_ex2_f32": (C.c_int, [_i64, _vp, C.POINTER(_dbl), _vp]),
}

_lib = None


def load():
    """Load libais_b200.so (once). Raises AisbError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AisbError(
            "libais_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ais_benchmarks_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale w.r.t. the header
        fn.restype = res
        fn.argtypes = args
    if lib.aisb_abi_version() != 1:
        raise AisbError("libais_b200.so ABI version %d, expected 1" % lib.aisb_abi_version())
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != AISB_OK:
        msg = load().aisb_last_error().decode("utf-8", "replace")
        raise AisbError("%s failed (status %d): %s" % (what or "libais_b200 call", rc, msg))
