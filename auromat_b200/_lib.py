"""ctypes binding of the C ABI declared in include/auromat_b200.h.

There is deliberately NO fallback: if the shared library is missing or cannot be loaded the
import of any compute path raises.  The reference's arrays are produced by CUDA kernels or
not at all.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AMT_LIB: development only -- A/B runs of differently compiled builds of the same sources on the GPU box
LIB_PATH = os.environ.get("AMT_LIB") or os.path.join(_HERE, "csrc", "libauromat_b200.so")

ABI_VERSION = 3          # AMT_ABI_VERSION of include/auromat_b200.h
AMT_OK, AMT_ERR_INVALID_ARGUMENT, AMT_ERR_UNSUPPORTED, AMT_ERR_CUDA, AMT_ERR_NO_DEVICE = range(5)
AMT_SIP_MAX_ORDER = 9
AMT_SIP_MAX_COEF = 55
AMT_U8, AMT_U16 = 0, 1
AMT_PRE_NONE, AMT_PRE_WRAP180, AMT_PRE_POLE = 0, 1, 2
AMT_MODEL_WCS, AMT_MODEL_ALLSKY = 0, 1

c_double_p = C.POINTER(C.c_double)


class AmtFrame(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("fast_center", C.c_int32), ("origin_inside", C.c_int32),
        ("crpix", C.c_double * 2), ("cd", C.c_double * 4), ("rot", C.c_double * 9),
        ("cam", C.c_double * 3), ("inv_axes", C.c_double * 3),
        ("m_geo", C.c_double * 9), ("m_sm", C.c_double * 9),
        ("wgs_a", C.c_double), ("wgs_b", C.c_double),
        ("sip_order_a", C.c_int32), ("sip_order_b", C.c_int32),
        ("sip_a", C.c_double * AMT_SIP_MAX_COEF), ("sip_b", C.c_double * AMT_SIP_MAX_COEF),
        ("model", C.c_int32), ("reserved", C.c_int32),
        ("allsky_xc", C.c_double), ("allsky_yc", C.c_double), ("allsky_k", C.c_double),
        ("allsky_rotation", C.c_double),
    ]


class AmtGeorefOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_lat_k", "d_lon_k", "d_mlat_k", "d_mlt_k",
        "d_lat_c", "d_lon_c", "d_mlat_c", "d_mlt_c", "d_elev_c", "d_valid_k", "d_valid_c")]


class AmtStats(C.Structure):
    _fields_ = [
        ("lat_min", C.c_double), ("lat_max", C.c_double),
        ("lon_min", C.c_double), ("lon_max", C.c_double),
        ("lon_min_pos", C.c_double), ("lon_max_neg", C.c_double),
        ("n_valid_corners", C.c_uint64), ("n_boundary_corners", C.c_uint64),
        ("n_valid_centers", C.c_uint64), ("n_ill_conditioned", C.c_uint64),
        ("pole_flags", C.c_uint64),
        ("row_min_c", C.c_int32), ("row_max_c", C.c_int32), ("col_min_c", C.c_int32), ("col_max_c", C.c_int32),
    ]


class AmtGrid(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("prerotate", C.c_int32), ("reserved", C.c_int32),
        ("lo_x", C.c_double), ("hi_x", C.c_double), ("step_x", C.c_double),
        ("lo_y", C.c_double), ("hi_y", C.c_double), ("step_y", C.c_double),
        ("round_x", C.c_double), ("round_y", C.c_double),
        ("altitude", C.c_double), ("wgs_a", C.c_double), ("wgs_b", C.c_double),
        ("rot", C.c_double * 9),
        ("side_scale", C.c_double),
    ]


class AmtGridInfo(C.Structure):
    _fields_ = [("n_lat", C.c_int32), ("n_lon", C.c_int32),
                ("lat_min_in_grid", C.c_double), ("lat_max_in_grid", C.c_double),
                ("lon_min_in_grid", C.c_double), ("lon_max_in_grid", C.c_double),
                ("lat_step", C.c_double), ("lon_step", C.c_double),
                ("lat_px_per_deg", C.c_double), ("lon_px_per_deg", C.c_double)]


AMT_PLAN_OK, AMT_PLAN_HOST, AMT_PLAN_EMPTY = 0, 1, 2


class AmtSeqSlot(C.Structure):
    _fields_ = [("planes", AmtGeorefOut), ("d_stats", C.c_void_p), ("h_stats", C.c_void_p), ("d_img", C.c_void_p)]


class AmtSeqJob(C.Structure):
    _fields_ = [("grid", C.POINTER(AmtGrid)), ("h_img", C.c_void_p), ("d_img", C.c_void_p),
                ("row0", C.c_int32), ("row1", C.c_int32), ("col0", C.c_int32), ("col1", C.c_int32),
                ("d_acc", C.c_void_p), ("d_out", C.c_void_p), ("h_out", C.c_void_p), ("out_bytes", C.c_size_t)]


# name -> (restype, argtypes); every symbol declared in include/auromat_b200.h
SIGNATURES = {
    "amt_last_error": (C.c_char_p, []),
    "amt_abi_version": (C.c_int, []),
    "amt_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "amt_ctx_destroy": (C.c_int, [C.c_void_p]),
    "amt_ctx_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "amt_ctx_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "amt_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "amt_measure_atomic_peak": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]),
    "amt_alloc_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "amt_free_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amt_alloc_pinned": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "amt_free_pinned": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amt_copy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_copy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "amt_copy_h2d_2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                  C.c_void_p]),
    "amt_memset_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]),
    "amt_stream_synchronize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "amt_georef": (C.c_int, [C.c_void_p, C.POINTER(AmtFrame), C.POINTER(AmtGeorefOut), C.c_void_p, C.c_void_p]),
    "amt_sanitize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(AmtGeorefOut), C.c_void_p]),
    "amt_valid_bits": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "amt_bbox_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int32, C.POINTER(AmtGrid), C.c_void_p, C.c_void_p]),
    "amt_apply_center_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double,
                                        C.POINTER(AmtGeorefOut), C.c_void_p]),
    "amt_rotate_coords": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(AmtGrid), C.c_void_p]),
    "amt_reproject": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_corner_means": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "amt_polygon_center_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int32, C.POINTER(AmtGrid), C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_plate_carree_coords": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_latlon_to_mlatmlt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double,
                                        C.c_double, c_double_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_sm_to_latlon": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, c_double_p, C.c_double, C.c_double,
                                   C.c_void_p]),
    "amt_bin_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_int32, C.c_size_t, C.POINTER(AmtGrid), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_cell_indices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(AmtGrid),
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_normalise": (C.c_int, [C.c_void_p, C.POINTER(AmtGrid), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "amt_bbox_stats_frame": (C.c_int, [C.c_void_p, C.POINTER(AmtFrame), C.c_void_p, C.c_void_p,
                                       C.POINTER(AmtGrid), C.c_void_p, C.c_void_p]),
    "amt_georef_fused": (C.c_int, [C.c_void_p, C.POINTER(AmtFrame), C.POINTER(AmtGeorefOut), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int32, C.c_int32, C.POINTER(AmtGrid), C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "amt_sip_distort": (C.c_int, [C.c_void_p, C.POINTER(AmtFrame), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "amt_plate_carree_resolution": (C.c_int, [C.c_double] * 5 + [c_double_p, c_double_p]),
    "amt_target_grid": (C.c_int, [C.c_double] * 6 + [C.POINTER(AmtGrid), C.POINTER(AmtGridInfo), C.POINTER(C.c_int32)]),
    "amt_side_scale": (C.c_double, [C.c_uint64]),
    "amt_sip_displacement_bound": (C.c_int, [C.POINTER(AmtFrame), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "amt_pole_pixels": (C.c_int, [C.POINTER(AmtFrame), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32)]),
    "amt_seq_plan": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.POINTER(AmtStats),
                               C.POINTER(AmtGrid), C.POINTER(AmtGridInfo), C.POINTER(C.c_int32)]),
    "amt_seq_output_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "amt_seq_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "amt_seq_destroy": (C.c_int, [C.c_void_p]),
    "amt_seq_set_slot": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AmtSeqSlot)]),
    "amt_seq_stage_a": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AmtFrame)]),
    "amt_seq_wait_stats": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AmtStats)]),
    "amt_seq_stage_b": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(AmtSeqJob)]),
    "amt_seq_wait_result": (C.c_int, [C.c_void_p, C.c_int32]),
    "amt_seq_h2d_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "amt_seq_trace": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double)]),
    "amt_georef_bin_fused": (C.c_int, [C.c_void_p, C.POINTER(AmtFrame), C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.POINTER(AmtGrid), C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
}

_lib = None


class AmtError(RuntimeError):
    pass


def load():
    """Load libauromat_b200.so (once).  Raises if it has not been built -- run
    `python -m auromat_b200.csrc.build` (or `__graft_entry__.build()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "auromat_b200: the CUDA extension %s is missing. Build it with "
            "`python -m auromat_b200.csrc.build`; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.amt_abi_version() != ABI_VERSION:
        raise ImportError("auromat_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(status: int):
    """Map the C status codes onto the exception types the reference raises
    (error conventions)."""
    if status == AMT_OK:
        return
    msg = load().amt_last_error().decode("utf-8", "replace")
    if status == AMT_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if status == AMT_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise AmtError(msg)
