"""opticalflow_ri_b200 -- B200 (sm_100a) implementation of OpticalFlow-RI's variational hot path.

Horn-Schunck Jacobi + Liu-Shen physics-based solve inside the coarse-to-fine pyramidal driver, as hand-written CUDA
kernels behind a C ABI (include/ofri.h, libofri.so) called through ctypes.  `opticalflow_ri_b200/src/` holds drop-in
modules with the reference's own module / class / function names; put that directory on sys.path where the
reference's scripts put `../src`.

There is no CPU implementation here: importing is cheap, but any computation needs libofri.so and a B200."""
import os

from ._lib import ALGO_EXTERNAL, ALGO_HS, ALGO_LS, ALGO_NONE, Algo, OfriError, Params, declared_symbols, lib  # noqa: F401
from .api import (Handle, LocalGroup, band_plan_host, default_handle, external_algo, farneback_params, fb_algo, lk_algo, lk_params, gaussian_taps, hs_algo, ls_algo, make_params, no_algo,  # noqa: F401
                  nccl_unique_id)

SRC_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "src")
__all__ = ["Handle", "LocalGroup", "nccl_unique_id", "band_plan_host", "default_handle", "make_params", "hs_algo", "ls_algo", "no_algo", "external_algo", "fb_algo", "farneback_params", "lk_algo", "lk_params", "gaussian_taps", "Params", "Algo",
           "OfriError", "lib", "declared_symbols", "SRC_DIR"]
