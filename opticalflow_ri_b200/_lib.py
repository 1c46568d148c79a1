"""ctypes binding of libofri.so (include/ofri.h).  No CPU fallback: if the CUDA library is missing or no B200 is
visible, loading / handle creation raises -- nothing in this package computes on the host."""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# OFRI_LIB: load another build of the same library (tools/phase_timing.py: libofri_phase.so); still CUDA-only
LIB_PATH = os.environ.get("OFRI_LIB") or os.path.join(HERE, "libofri.so")
HEADER = os.path.join(HERE, "..", "include", "ofri.h")

OFRI_MAX_GAUSS_TAPS = 129
OFRI_MAX_ALPHAS = 64
ALGO_NONE, ALGO_HS, ALGO_LS, ALGO_EXTERNAL, ALGO_FB, ALGO_LK = -1, 0, 1, 2, 3, 4
OFRI_FB_MAX_HALF, OFRI_FB_MAX_LEVELS = 64, 12

OK = 0
ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM, ERR_ALPHAS, ERR_FILTER_OPT, ERR_TOO_SMALL, ERR_UNSUPPORTED, ERR_COMM, \
    ERR_INDEX, ERR_CALLBACK = -1, -2, -3, -4, -5, -6, -7, -8, -9, -10, -11


class OfriError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libofri error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class Algo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("hs_niter", C.c_int32), ("n_alphas", C.c_int32), ("ls_maxiter", C.c_int32),
                ("alphas", C.c_float * OFRI_MAX_ALPHAS), ("ls_h", C.c_float), ("reserved_", C.c_float),
                ("ls_tol", C.c_double)]


class Params(C.Structure):
    _fields_ = [("size", C.c_uint32), ("pyramid_levels", C.c_int32), ("k_levels", C.c_int32), ("warping", C.c_int32),
                ("bilinear", C.c_int32), ("intermediate_scaling", C.c_int32), ("final_scaling", C.c_int32),
                ("n_taps_main", C.c_int32), ("n_taps_opt", C.c_int32), ("refilter_k", C.c_int32),
                ("taps_main", C.c_float * OFRI_MAX_GAUSS_TAPS), ("taps_opt", C.c_float * OFRI_MAX_GAUSS_TAPS),
                ("main_algo", Algo), ("opt_algo", Algo), ("n_taps_lsw", C.c_int32),
                ("taps_lsw", C.c_float * OFRI_MAX_GAUSS_TAPS)]


class FarnebackParams(C.Structure):
    _fields_ = [("size", C.c_uint32), ("window_size", C.c_int32), ("n_iters", C.c_int32), ("poly_n", C.c_int32),
                ("use_gaussian", C.c_int32), ("extra_levels", C.c_int32), ("pyr_scale", C.c_float),
                ("g", C.c_float * 8), ("xg", C.c_float * 8), ("xxg", C.c_float * 8), ("ig", C.c_float * 4),
                ("win_kernel", C.c_float * (OFRI_FB_MAX_HALF + 1)), ("n_blur", C.c_int32 * OFRI_FB_MAX_LEVELS),
                ("blur_kernel", (C.c_float * (OFRI_FB_MAX_HALF + 1)) * OFRI_FB_MAX_LEVELS)]


class LkParams(C.Structure):
    _fields_ = [("size", C.c_uint32), ("n_iters", C.c_int32), ("half_window", C.c_int32), ("asym", C.c_int32 * 4)]


class Band(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("own0", C.c_int32), ("own1", C.c_int32),
                ("in0", C.c_int32), ("in1", C.c_int32), ("ghost", C.c_int32), ("exchange", C.c_int32)]


_fp = C.POINTER(C.c_float)
_H = C.c_void_p
# ofri_adapter_fn: int fn(void* user, int which, int call_index, const float* im1, const float* im2, float* U, float* V,
#                         int H, int W, float* err)
ADAPTER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, _fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp)
_SIGNATURES = {
    "ofri_abi_version": (C.c_int, []),
    "ofri_device_count": (C.c_int, []),
    "ofri_create": (C.c_int, [C.c_int, C.POINTER(_H)]),
    "ofri_destroy": (C.c_int, [_H]),
    "ofri_last_error": (C.c_char_p, [_H]),
    "ofri_set_stream": (C.c_int, [_H, C.c_void_p]),
    "ofri_synchronize": (C.c_int, [_H]),
    "ofri_set_option": (C.c_int, [_H, C.c_char_p, C.c_int]),
    "ofri_get_option": (C.c_int, [_H, C.c_char_p, C.POINTER(C.c_int)]),
    "ofri_launch_count": (C.c_int64, [_H]),
    "ofri_stage_timings": (C.c_int, [_H, C.POINTER(C.c_char_p), _fp, C.c_int]),
    "ofri_debug_phase_read": (C.c_int, [_H, C.c_int, C.POINTER(C.c_ulonglong)]),
    "ofri_pyramidal_flow": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Params),
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "ofri_farneback_compute": (C.c_int, [_H, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(FarnebackParams), _fp, _fp]),
    "ofri_set_farneback": (C.c_int, [_H, C.POINTER(FarnebackParams)]),
    "ofri_lk_compute": (C.c_int, [_H, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.POINTER(LkParams), _fp, _fp]),
    "ofri_set_lk": (C.c_int, [_H, C.POINTER(LkParams)]),
    "ofri_resize_bilinear": (C.c_int, [_H, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "ofri_host_alloc": (C.c_int, [_H, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ofri_host_free": (C.c_int, [_H, C.c_void_p]),
    "ofri_pyramidal_flow_external": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Params), ADAPTER_FN,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ofri_pyramidal_flow_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Params),
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "ofri_hs_compute": (C.c_int, [_H, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _fp, _fp, _fp]),
    "ofri_ls_compute": (C.c_int, [_H, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double,
                                  _fp, _fp, _fp, C.POINTER(C.c_int32)]),
    "ofri_gauss_px": (C.c_int, [_H, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp]),
    "ofri_gaussian_taps": (C.c_int, [C.c_double, C.c_int, _fp]),
    "ofri_resize_bicubic": (C.c_int, [_H, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "ofri_level_size": (C.c_int, [C.c_int, C.c_double]),
    "ofri_spline_upsample": (C.c_int, [_H, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "ofri_warp_bilinear": (C.c_int, [_H, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "ofri_warp_pair": (C.c_int, [_H, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp]),
    "ofri_liu_shen_warp": (C.c_int, [_H, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, C.c_int, _fp]),
    "ofri_hs_derivatives": (C.c_int, [_H, _fp, _fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp]),
    "ofri_hs_iterate": (C.c_int, [_H, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _fp, _fp]),
    "ofri_ls_coefficients": (C.c_int, [_H, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "ofri_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ofri_comm_init_nccl": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    "ofri_local_group_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ofri_local_group_destroy": (C.c_int, [C.c_void_p]),
    "ofri_local_group_abort": (C.c_int, [C.c_void_p]),
    "ofri_comm_init_local": (C.c_int, [_H, C.c_void_p, C.c_int]),
    "ofri_comm_destroy": (C.c_int, [_H]),
    "ofri_band_plan": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(Params), C.c_int, C.c_int, C.POINTER(Band)]),
    "ofri_band_plan_host": (C.c_int, [C.c_int, C.c_int, C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(Band)]),
    "ofri_pyramidal_flow_banded_dev": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(Params),
                                                 C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


def declared_symbols():
    """Every function include/ofri.h declares (used by the CPU test that checks the exports)."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"OFRI_API\s+[\w\s\*]+?\b(ofri_\w+)\s*\(", f.read())))


def lib():
    """Load libofri.so.  Raises if it has not been built -- there is no pure-Python / CPU implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libofri.so not found at %s: build it with `python -m opticalflow_ri_b200.build` "
                              "(CUDA-only library, no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)     # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.ofri_abi_version() != 1:
            raise ImportError("libofri.so ABI version %d, binding expects 1" % L.ofri_abi_version())
        if C.sizeof(Params) == 0:
            raise ImportError("bad Params layout")
        _lib = L
    return _lib
