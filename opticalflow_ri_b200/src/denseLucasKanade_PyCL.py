"""Drop-in for the reference's src/denseLucasKanade_PyCL.py: the same class name, constructor parameters and adapter
protocol, computing on the B200 through libofri.so (csrc/ofri_lk.cu) instead of PyOpenCL.

Mirrors denseLucasKanade_PyCL.py:33-182: constructor (platformID, deviceID, Niter, halfWindow,
provideGenericPyramidalDefaults, enableVorticityEnhancement -- the two OpenCL ids are accepted and ignored),
evaluateVorticityEnhancement (:75-92, host side: it reduces the incoming flow to four 0/1 switches),
compute(im1, im2, U, V) -> (U, V, True) (:113-169; the reference returns its `calcErr` flag as the third value) and the
generic-pyramid defaults {'warping': False, 'intermediateScaling': True, 'scaling': False} (:177-182).  Inside
genericPyramidalOpticalFlow the adapter runs natively with the rest of the driver unless the vorticity enhancement is
on (its switches depend on the flow of every call, so the driver then hands each level to compute() below)."""
import numpy as np

import _native
from _native import ofri


class denseLucasKanade_PyCl(object):
    def __init__(self, platformID=0, deviceID=0, Niter=5, halfWindow=13, provideGenericPyramidalDefaults=True,
                 enableVorticityEnhancement=False):
        self.provideGenericPyramidalDefaults = provideGenericPyramidalDefaults
        self.enableVorticityEnhancement = enableVorticityEnhancement
        self.windowHalfWidth = self.windowHalfHeight = int(halfWindow)
        self.windowWidth = self.windowHeight = 2 * int(halfWindow) + 1
        self.Niter = Niter

    @property
    def _ofri_native_kind(self):
        return None if self.enableVorticityEnhancement else "LK"

    def evaluateVorticityEnhancement(self, U, V):
        """[left, right, top, bottom] window switches from the sign of the mean vorticity of (U, V)."""
        if not self.enableVorticityEnhancement:
            return [0, 0, 0, 0]
        U = np.asarray(U, np.float32)
        V = np.asarray(V, np.float32)
        # central differences with scipy's 'reflect' rule (the sample beyond the border repeats the border sample)
        Vp = np.pad(V, ((0, 0), (1, 1)), mode='symmetric')
        Up = np.pad(U, ((1, 1), (0, 0)), mode='symmetric')
        half = np.float32(0.5)
        # scipy.ndimage.convolve flips the kernel: D = [[0,-1,0],[0,0,0],[0,1,0]] / 2 gives (x[i-1] - x[i+1]) / 2
        Dv = Vp[:, :-2] * half + Vp[:, 2:] * -half
        Du = Up[:-2, :] * half + Up[2:, :] * -half
        m = np.mean(Dv - Du)
        if m < -2e-3:
            return [0, 1, 0, 1]
        if m > 2e-3:
            return [1, 0, 0, 1]
        return [0, 0, 0, 0]

    def native_params(self, asym=(0, 0, 0, 0)):
        return ofri.lk_params(self.Niter, self.windowHalfWidth, asym)

    def compute(self, im1, im2, U, V):
        assert np.shape(im1) == np.shape(im2) == np.shape(U) == np.shape(V)
        asym = self.evaluateVorticityEnhancement(U, V)
        # a handle of its own: compute() may be called back from inside the driver's native call on the default handle
        Un, Vn = _native.aux_handle().lk_compute(im1, im2, U, V, self.native_params(asym))
        return Un, Vn, True

    def getAlgoName(self):
        return 'OpenCL Dense LK'

    def hasGenericPyramidalDefaults(self):
        return self.provideGenericPyramidalDefaults

    def getGenericPyramidalDefaults(self):
        return {'warping': False, 'intermediateScaling': True, 'scaling': False}
