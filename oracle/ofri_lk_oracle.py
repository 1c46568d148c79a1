"""ctypes front end of oracle/ofri_lk_oracle.c (dense Lucas-Kanade restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__ and bench.py's CPU legs may import this file.  The C file is compiled on first use (gcc,
`-ffp-contract=off`) into oracle/_build/ (git-ignored).  See the C file's header for what is and is not pinned.
Also restates the adapter's host side: the asymmetric-window switch of denseLucasKanade_PyCL.py:75-92 (LK:line) and the
generic-pyramid defaults (LK:177-182)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ofri_lk_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libofri_lk_oracle.so")
_lib = None


def build():
    if not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", _SRC, "-o", _OUT, "-lm"], check=True)
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        _lib.ofri_lk_oracle.restype = C.c_int
        _lib.ofri_lk_oracle.argtypes = [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, ip]
    return _lib


def vorticity_switch(U, V, enabled):
    """evaluateVorticityEnhancement (LK:75-92): [left, right, top, bottom] from the sign of the mean vorticity."""
    if not enabled:
        return [0, 0, 0, 0]
    from scipy.ndimage import convolve
    D = np.array([[0, -1, 0], [0, 0, 0], [0, 1, 0]], dtype=np.float32) * np.float32(0.5)
    omega = convolve(V, D.T, mode='reflect') - convolve(U, D, mode='reflect')
    m = np.mean(omega)
    if m < -2e-3:
        return [0, 1, 0, 1]
    if m > 2e-3:
        return [1, 0, 0, 1]
    return [0, 0, 0, 0]


def lk_compute(im1, im2, U, V, n_iters=5, half_window=13, asym=(0, 0, 0, 0)):
    im1 = np.ascontiguousarray(im1, np.float32)
    im2 = np.ascontiguousarray(im2, np.float32)
    u = np.array(U, np.float32, order="C")
    v = np.array(V, np.float32, order="C")
    H, W = im1.shape
    fp = C.POINTER(C.c_float)
    a = (C.c_int * 4)(*[int(x) for x in asym])
    rc = lib().ofri_lk_oracle(im1.ctypes.data_as(fp), im2.ctypes.data_as(fp), u.ctypes.data_as(fp), v.ctypes.data_as(fp),
                              H, W, int(half_window), int(n_iters), a)
    assert rc == 0
    return u, v


class LKParams:
    """Oracle-side stand-in for denseLucasKanade_PyCl (LK:33-182): same constructor defaults."""
    name = "LK"

    def __init__(self, Niter=5, halfWindow=13, provide_defaults=True, enableVorticityEnhancement=False):
        self.Niter, self.halfWindow = Niter, halfWindow
        self.provide_defaults, self.vorticity = provide_defaults, enableVorticityEnhancement

    def defaults(self):
        return {"warping": False, "intermediateScaling": True, "scaling": False} if self.provide_defaults else None

    def compute(self, im1, im2, U, V):
        asym = vorticity_switch(np.asarray(U, np.float32), np.asarray(V, np.float32), self.vorticity)
        u, v = lk_compute(im1, im2, U, V, self.Niter, self.halfWindow, asym)
        return u, v, True          # the reference returns `calcErr` (= True) as its third value (LK:169)
