"""Drop-in for the reference's src/HornSchunck.py (same class / function names and argument meaning), computing on
the B200 through libofri.so.  Reference behaviour mirrored: HornSchunck.py:29-50 (adapter), 73-105 (HS), 107-127
(computeDerivatives).  No numba / scipy on this path and no CPU fallback."""
import numpy as np

import _native


class HSOpticalFlowAlgoAdapter(object):
    _ofri_native_kind = "HS"

    def __init__(self, alphas, Niter, provideGenericPyramidalDefaults=True):
        self.provideGenericPyramidalDefaults = provideGenericPyramidalDefaults
        self.alphas = alphas
        self.Niter = Niter

    def compute(self, im1, im2, U, V):
        alpha = self.alphas.pop()          # last alpha first; IndexError when the list runs out (reference :36)
        return HS(im1, im2, alpha, self.Niter, U, V)

    def getAlgoName(self):
        return 'Horn-Schunck'

    def hasGenericPyramidalDefaults(self):
        return self.provideGenericPyramidalDefaults

    def getGenericPyramidalDefaults(self):
        return {'warping': True, 'biLinear': True, 'scaling': True}


def HS(im2, im1, alpha, Niter, U, V):
    """Same (historically swapped) parameter names as the reference: the adapter calls HS(frame1, frame2, ...).
    Returns (U, V, total_error) after exactly Niter Jacobi sweeps."""
    frame1, frame2 = im2, im1
    Un, Vn, err = _native.handle().hs_compute(frame1, frame2, U, V, alpha, Niter)
    return Un, Vn, err


def computeDerivatives(im1, im2):
    """fx, fy, ft with the reference's argument order: fx = g(im1)+g(im2), ft = box(im2) - box(im1)."""
    return _native.handle().hs_derivatives(im2, im1)
