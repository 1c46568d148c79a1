"""Drop-in for the reference's src/gaussian_filter.py: same functions, the separable filter runs on the B200.
Coefficients are generated on the host exactly as gaussian_filter.py:47-52 does (bit-identical float32 taps)."""
import numpy as np

import _native


def prepareGaussianKernel(sigma, kernelSizePx):
    return _native.ofri.gaussian_taps(sigma, kernelSizePx)


def convolveSeparableFilter(kernel, image):
    """Rows then columns with the reference's edge padding (gaussian_filter.py:54-85); like the reference it
    overwrites `image` when that is a float32 array and returns it."""
    out = _native.handle().gauss_px(image, kernel)
    if isinstance(image, np.ndarray) and image.dtype == np.float32 and image.flags.writeable:
        image[...] = out
        return image
    return out


def gaussian_filter(image, sigma, truncate):
    kernelSizePx = 2 * int(truncate * sigma + 0.5) + 1
    return convolveSeparableFilter(prepareGaussianKernel(sigma, kernelSizePx), image)


def gaussian_filterPx(image, sigma, kernelSizePx):
    return convolveSeparableFilter(prepareGaussianKernel(sigma, kernelSizePx), image)
