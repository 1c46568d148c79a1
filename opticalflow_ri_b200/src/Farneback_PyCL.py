"""Drop-in for the reference's src/Farneback_PyCL.py: the same class name, constructor parameters and adapter protocol,
computing on the B200 through libofri.so (csrc/ofri_farneback.cu) instead of PyOpenCL.

Mirrors Farneback_PyCL.py:65-616: constructor (windowSize, Niters, polyN, polySigma, useGaussian, pyrScale,
pyramidalLevels, platformID, deviceID, provideGenericPyramidalDefaults -- the two OpenCL ids are accepted and ignored),
the host-side coefficient tables (FarnebackPrepareGaussian :124-176, setGaussianBlurKernel :199-207 through this
package's bit-identical getGaussianKernelBitExact), compute(im1, im2, U, V) -> (U, V, 'Unknown') (:462-604) and the
generic-pyramid defaults {'warping': False, 'scaling': True} (:612-616).  Inside genericPyramidalOpticalFlow the adapter
runs natively with the rest of the driver (no host round trips per level)."""
import numpy as np

import _native
from _native import ofri
from GaussianKernelBitExact import getGaussianKernelBitExact


def imresize(im, res):
    """Pillow-BILINEAR resample to res = (width, height) (reference :61-62) on the GPU."""
    return _native.handle().resize_bilinear(im, int(res[1]), int(res[0]))


class Farneback_PyCL(object):
    _ofri_native_kind = "FB"

    def __init__(self, windowSize=33, Niters=5, polyN=7, polySigma=1.5, useGaussian=True, pyrScale=0.5, pyramidalLevels=1,
                 platformID=0, deviceID=0, provideGenericPyramidalDefaults=True):
        assert pyramidalLevels >= 1, 'Pyramidal levels must be greater or equal than 1'
        self.useDouble = False
        self.windowSize = windowSize
        self.numIters = Niters
        self.polyN = int(polyN)
        self.polySigma = polySigma
        self.useGaussianFilter = useGaussian
        self.pyramidalLevels = pyramidalLevels - 1
        self.fastPyramids = False
        self.pyrScale = pyrScale
        self.provideGenericPyramidalDefaults = provideGenericPyramidalDefaults
        if windowSize & 1 == 0:
            raise Exception('windowSize must be an odd value')

    # ---- host-side tables, computed exactly as the reference does --------------------------------------------------------
    def FarnebackPrepareGaussian(self):
        n = self.polyN
        sigma = self.polySigma
        if sigma < 1.19209289550781250000000000000000000e-7:
            sigma = n * 0.3
        g = np.zeros([2 * n + 1], dtype=np.float32)
        xg = np.zeros([2 * n + 1], dtype=np.float32)
        xxg = np.zeros([2 * n + 1], dtype=np.float32)
        s = np.float64(0.0)
        for x in range(-n, n + 1):
            g[x + n] = np.exp(-x * x / (2 * sigma * sigma))
            s += g[x + n]
        s = 1.0 / s
        for x in range(-n, n + 1):
            g[x + n] = np.float32(g[x + n] * s)
            xg[x + n] = np.float32(x * g[x + n])
            xxg[x + n] = np.float32(x * x * g[x + n])
        G = np.zeros((6, 6), np.float64)
        for y in range(-n, n + 1):
            for x in range(-n, n + 1):
                G[0, 0] += g[y + n] * g[x + n]
                G[1, 1] += g[y + n] * g[x + n] * x * x
                G[3, 3] += g[y + n] * g[x + n] * x * x * x * x
                G[5, 5] += g[y + n] * g[x + n] * x * x * y * y
        G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
        G[4, 4] = G[3, 3]
        G[3, 4] = G[4, 3] = G[5, 5]
        invG = np.linalg.inv(G)
        return g, xg, xxg, invG[1, 1], invG[0, 3], invG[3, 3], invG[5, 5]

    def getGaussianKernel(self, n, sigma):
        _, kernel_bitexact = getGaussianKernelBitExact(n, sigma)
        return np.float32(kernel_bitexact).reshape([1, n])

    def _kernel_half(self, size, sigma):
        g = self.getGaussianKernel(size, sigma)
        return g[0, int(size / 2):].copy()

    def native_params(self):
        """ofri_farneback_params of this adapter (tables for every internal pyramid level it may use); rebuilt only when
        a constructor attribute has changed (the Decimal arithmetic of the window kernel takes milliseconds)."""
        key = (self.windowSize, self.numIters, self.polyN, self.polySigma, bool(self.useGaussianFilter),
               self.pyramidalLevels, self.pyrScale)
        if getattr(self, '_params_key', None) != key:
            self._params, self._params_key = self._build_native_params(), key
        return self._params

    def _build_native_params(self):
        n = self.polyN
        g, xg, xxg, ig11, ig03, ig33, ig55 = self.FarnebackPrepareGaussian()
        blur = []
        for k in range(self.pyramidalLevels + 1):
            scale = 1.0
            for _ in range(k):
                scale *= self.pyrScale
            sigma = (1.0 / scale - 1.0) * 0.5
            smoothSize = max(int(round(sigma * 5)) | 1, 3)
            blur.append(self._kernel_half(smoothSize, sigma)[:int(smoothSize / 2) + 1])
        win = self._kernel_half(self.windowSize, self.windowSize / 2 * 0.3) if self.useGaussianFilter else np.zeros(1, np.float32)
        return ofri.farneback_params(self.windowSize, self.numIters, n, self.useGaussianFilter, self.pyramidalLevels,
                                     self.pyrScale, g[n:], xg[n:], xxg[n:],
                                     np.float32([ig11, ig03, ig33, ig55]), win, blur)

    # ---- adapter protocol ----------------------------------------------------------------------------------------------------
    def compute(self, im1, im2, U, V):
        assert self.polyN == 5 or self.polyN == 7
        assert im1.shape == im2.shape and self.pyrScale < 1
        assert U.shape == im1.shape and V.shape == im1.shape
        Un, Vn = _native.handle().farneback_compute(im1, im2, U, V, self.native_params())
        return Un, Vn, 'Unknown'

    def getAlgoName(self):
        return 'Farneback CL'

    def hasGenericPyramidalDefaults(self):
        return self.provideGenericPyramidalDefaults

    def getGenericPyramidalDefaults(self):
        return {'warping': False, 'scaling': True}
