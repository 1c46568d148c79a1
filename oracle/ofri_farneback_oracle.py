"""CPU restatement (numpy) of the reference's Farneback adapter -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product (opticalflow_ri_b200)
never does.  Follows `/root/reference/src/Farneback_PyCL.py` (host logic, cited as FB:line) and the six OpenCL kernels of
`/root/reference/src/optical_flow_farneback.cl` (cited as CL:line), float32 arithmetic with every multiply and add
rounded separately, in the order the kernels write them.

PARITY PINNING: the reference's own implementation of this adapter needs an OpenCL runtime (pyopencl), which this image
does not have, so its outputs cannot be generated here: **parity unpinned** against the reference itself.  What the tests
do instead: (1) every coefficient table against the reference's pure-Python generators (FarnebackPrepareGaussian is
restated; getGaussianKernelBitExact is shared with the pinned oracle), (2) the whole compute() against OpenCV's CPU
implementation of the same algorithm (cv2.calcOpticalFlowFarneback, which the reference's kernels port; agreement to
< 3e-6 px at the fixed point of the iteration in the image interior, tests/test_farneback_cpu.py), (3) the CUDA path against this file on seeded inputs.
"""
import numpy as np

import ofri_oracle as O

F32 = np.float32


# ---- coefficient tables (FB:124-207) ---------------------------------------------------------------------------------
def prepare_gaussian(poly_n, poly_sigma):
    """FarnebackPrepareGaussian (FB:124-176): g, xg, xxg (float32, 2n+1) and ig11, ig03, ig33, ig55 (float64)."""
    n = int(poly_n)
    sigma = float(poly_sigma)
    if sigma < 1.19209289550781250000000000000000000e-7:
        sigma = n * 0.3
    g = np.zeros(2 * n + 1, F32)
    xg = np.zeros(2 * n + 1, F32)
    xxg = np.zeros(2 * n + 1, F32)
    s = np.float64(0.0)
    for x in range(-n, n + 1):
        g[x + n] = np.exp(-x * x / (2 * sigma * sigma))
        s += g[x + n]
    s = 1.0 / s
    for x in range(-n, n + 1):
        g[x + n] = F32(g[x + n] * s)
        xg[x + n] = F32(x * g[x + n])
        xxg[x + n] = F32(x * x * g[x + n])
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


def blur_kernel_half(ksize, sigma):
    """setGaussianBlurKernel (FB:199-207): the centre and right half of getGaussianKernelBitExact(ksize, sigma) as f32."""
    _, k = O.gaussian_kernel_bit_exact(int(ksize), sigma)
    k = np.asarray(k, np.float64).astype(F32)
    return k[int(ksize / 2):].copy()


# ---- border index rules of the kernels (CL:134-157) ----------------------------------------------------------------------
def _idx_low(i, last):
    return np.abs(i) % (last + 1)


def _idx_high(i, last):
    return np.abs(last - np.abs(last - i)) % (last + 1)


def _idx_col(i, last):
    return _idx_low(_idx_high(i, last), last)


# ---- kernels ------------------------------------------------------------------------------------------------------------------
def gaussian_blur(src, gk):
    """gaussianBlur (CL:159-196): vertical pass into the row cache, then horizontal pass; reflect-101 style indices."""
    src = np.asarray(src, F32)
    rows, cols = src.shape
    kh = len(gk) - 1
    ys = np.arange(rows)
    xs = _idx_col(np.arange(-kh, cols + kh), cols - 1)
    ext = src[:, xs]                                              # columns x - kh .. x + kh as the kernel addresses them
    vert = ext * gk[0]
    for j in range(1, kh + 1):
        lo = _idx_low(ys - j, rows - 1)
        hi = _idx_high(ys + j, rows - 1)
        vert = vert + (ext[lo] + ext[hi]) * gk[j]
    vert = vert.astype(F32)
    c = np.arange(cols) + kh
    res = vert[:, c] * gk[0]
    for i in range(1, kh + 1):
        res = res + (vert[:, c - i] + vert[:, c + i]) * gk[i]
    return res.astype(F32)


def _blur5(M, gk, box):
    """gaussianBlur5 (CL:198-254) / boxFilter5 (CL:350-406) on five planes M[5][rows][cols]."""
    M = np.asarray(M, F32)
    _, rows, cols = M.shape
    kh = int(box) if box is not None else len(gk) - 1
    ys = np.arange(rows)
    if box is None:
        xs = _idx_col(np.arange(-kh, cols + kh), cols - 1)
    else:
        xs = np.clip(np.arange(-kh, cols + kh), 0, cols - 1)
    ext = M[:, :, xs]
    if box is None:
        vert = ext * gk[0]
        for j in range(1, kh + 1):
            vert = vert + (ext[:, _idx_low(ys - j, rows - 1)] + ext[:, _idx_high(ys + j, rows - 1)]) * gk[j]
    else:
        vert = ext.copy()
        for j in range(1, kh + 1):
            vert = vert + (ext[:, np.maximum(ys - j, 0)] + ext[:, np.minimum(ys + j, rows - 1)])
    vert = vert.astype(F32)
    c = np.arange(cols) + kh
    if box is None:
        res = vert[:, :, c] * gk[0]
        for i in range(1, kh + 1):
            res = res + (vert[:, :, c - i] + vert[:, :, c + i]) * gk[i]
    else:
        res = vert[:, :, c]
        for i in range(1, kh + 1):
            res = res + (vert[:, :, c - i] + vert[:, :, c + i])
        res = res * F32(F32(1.0) / F32((1 + 2 * kh) * (1 + 2 * kh)))
    return res.astype(F32)


def polynomial_expansion(src, g, xg, xxg, ig):
    """polynomialExpansion (CL:72-132): g / xg / xxg = the halves [0..n] of the tables; ig = float32[4]."""
    src = np.asarray(src, F32)
    rows, cols = src.shape
    n = len(g) - 1
    ys = np.arange(rows)
    xs = np.clip(np.arange(-n, cols + n), 0, cols - 1)            # xWarped (CL:90)
    ext = src[:, xs]
    r0 = ext * g[0]
    r1 = np.zeros_like(ext)
    r2 = np.zeros_like(ext)
    for k in range(1, n + 1):
        t0 = ext[np.maximum(ys - k, 0)]
        t1 = ext[np.minimum(ys + k, rows - 1)]
        r0 = r0 + g[k] * (t0 + t1)
        r1 = r1 + xg[k] * (t1 - t0)
        r2 = r2 + xxg[k] * (t0 + t1)
    r0, r1, r2 = r0.astype(F32), r1.astype(F32), r2.astype(F32)
    c = np.arange(cols) + n
    b1 = g[0] * r0[:, c]
    b3 = g[0] * r1[:, c]
    b5 = g[0] * r2[:, c]
    b2 = np.zeros((rows, cols), F32)
    b4 = np.zeros((rows, cols), F32)
    b6 = np.zeros((rows, cols), F32)
    for k in range(1, n + 1):
        b1 = b1 + (r0[:, c + k] + r0[:, c - k]) * g[k]
        b4 = b4 + (r0[:, c + k] + r0[:, c - k]) * xxg[k]
        b2 = b2 + (r0[:, c + k] - r0[:, c - k]) * xg[k]
        b3 = b3 + (r1[:, c + k] + r1[:, c - k]) * g[k]
        b6 = b6 + (r1[:, c + k] - r1[:, c - k]) * xg[k]
        b5 = b5 + (r2[:, c + k] + r2[:, c - k]) * g[k]
    out = np.empty((5, rows, cols), F32)
    out[0] = b3 * ig[0]
    out[1] = b2 * ig[0]
    out[2] = b1 * ig[1] + b5 * ig[2]
    out[3] = b1 * ig[1] + b4 * ig[2]
    out[4] = b6 * ig[3]
    return out


_BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472, 1.0], F32)       # CL:255 (BORDER_SIZE = 5)


def update_matrices(flowx, flowy, R0, R1):
    """updateMatrices (CL:256-348)."""
    flowx = np.asarray(flowx, F32)
    flowy = np.asarray(flowy, F32)
    rows, cols = flowx.shape
    yy, xx = np.mgrid[0:rows, 0:cols]
    dx, dy = flowx, flowy
    fx = (xx.astype(F32) + dx).astype(F32)
    fy = (yy.astype(F32) + dy).astype(F32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(F32)).astype(F32)
    fy = (fy - y1.astype(F32)).astype(F32)
    inside = (x1 >= 0) & (y1 >= 0) & (x1 < cols - 1) & (y1 < rows - 1)
    xc = np.clip(x1, 0, cols - 2)
    yc = np.clip(y1, 0, rows - 2)
    one = F32(1.0)
    a00 = (one - fx) * (one - fy)
    a01 = fx * (one - fy)
    a10 = (one - fx) * fy
    a11 = fx * fy

    def samp(p):
        return (a00 * R1[p][yc, xc] + a01 * R1[p][yc, xc + 1] + a10 * R1[p][yc + 1, xc] + a11 * R1[p][yc + 1, xc + 1]).astype(F32)

    r2i, r3i, r4i, r5i, r6i = (samp(p) for p in range(5))
    r4 = np.where(inside, (R0[2] + r4i) * F32(0.5), R0[2]).astype(F32)
    r5 = np.where(inside, (R0[3] + r5i) * F32(0.5), R0[3]).astype(F32)
    r6 = np.where(inside, (R0[4] + r6i) * F32(0.25), R0[4] * F32(0.5)).astype(F32)
    r2 = ((R0[0] - np.where(inside, r2i, F32(0))) * F32(0.5)).astype(F32)
    r3 = ((R0[1] - np.where(inside, r3i, F32(0))) * F32(0.5)).astype(F32)
    r2 = (r2 + (r4 * dy + r6 * dx)).astype(F32)                # r2 += r4*dy + r6*dx: the right-hand side is summed first
    r3 = (r3 + (r6 * dy + r5 * dx)).astype(F32)
    bs = 5
    scale = (_BORDER[np.minimum(xx, bs)] * _BORDER[np.minimum(yy, bs)] * _BORDER[np.minimum(cols - xx - 1, bs)] *
             _BORDER[np.minimum(rows - yy - 1, bs)]).astype(F32)
    r2, r3, r4, r5, r6 = (r * scale for r in (r2, r3, r4, r5, r6))
    M = np.empty((5, rows, cols), F32)
    M[0] = r4 * r4 + r6 * r6
    M[1] = (r4 + r5) * r6
    M[2] = r5 * r5 + r6 * r6
    M[3] = r4 * r2 + r6 * r3
    M[4] = r6 * r2 + r5 * r3
    return M


def update_flow(M):
    """updateFlow (CL:408-429)."""
    g11, g12, g22, h1, h2 = (np.asarray(M[i], F32) for i in range(5))
    det_inv = (F32(1.0) / (g11 * g22 - g12 * g12 + F32(1e-3))).astype(F32)
    return ((g11 * h2 - g12 * h1) * det_inv).astype(F32), ((g22 * h1 - g12 * h2) * det_inv).astype(F32)


def imresize_bilinear(im, w, h):
    """FB:61-62: Pillow BILINEAR (antialiased triangle filter) through the oracle's Pillow resampler."""
    return O.imresize_filter(np.asarray(im, F32), int(w), int(h), "bilinear")


# ---- the adapter (FB:462-604) -------------------------------------------------------------------------------------------------
class FBParams:
    """Oracle-side stand-in for Farneback_PyCL (FB:65-616): same constructor defaults."""
    name = "FB"

    def __init__(self, windowSize=33, Niters=5, polyN=7, polySigma=1.5, useGaussian=True, pyrScale=0.5, pyramidalLevels=1,
                 provide_defaults=True):
        if windowSize & 1 == 0:
            raise Exception("windowSize must be an odd value")
        self.windowSize, self.numIters, self.polyN, self.polySigma = windowSize, Niters, int(polyN), polySigma
        self.useGaussianFilter, self.pyrScale, self.pyramidalLevels = useGaussian, pyrScale, pyramidalLevels - 1
        self.provide_defaults = provide_defaults

    def defaults(self):
        return {"warping": False, "scaling": True} if self.provide_defaults else None

    def compute(self, im1, im2, U, V):
        assert self.polyN in (5, 7) and im1.shape == im2.shape and self.pyrScale < 1
        im1 = np.asarray(im1, F32)
        im2 = np.asarray(im2, F32)
        size = im1.shape
        min_size = 32
        scale = 1
        levels = 0
        while levels < self.pyramidalLevels:                      # FB:483-489
            scale *= self.pyrScale
            if size[1] * scale < min_size or size[0] * scale < min_size:
                break
            levels += 1
        g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(self.polyN, self.polySigma)
        n = self.polyN
        gh, xgh, xxgh = g[n:].copy(), xg[n:].copy(), xxg[n:].copy()
        ig = np.array([ig11, ig03, ig33, ig55], np.float64).astype(F32)
        prev = None
        cur_x = cur_y = None
        for k in range(levels, -1, -1):                           # FB:512
            scale = 1.0
            for _ in range(k):
                scale *= self.pyrScale
            sigma = (1.0 / scale - 1.0) * 0.5
            smooth = max(int(round(sigma * 5)) | 1, 3)
            width = int(round(size[1] * scale))
            height = int(round(size[0] * scale))
            if prev is None:
                cur_x = (imresize_bilinear(U, width, height) * F32(scale)).astype(F32)
                cur_y = (imresize_bilinear(V, width, height) * F32(scale)).astype(F32)
            else:
                cur_x = (imresize_bilinear(prev[0], width, height) * F32(1.0 / self.pyrScale)).astype(F32)
                cur_y = (imresize_bilinear(prev[1], width, height) * F32(1.0 / self.pyrScale)).astype(F32)
            gk = blur_kernel_half(smooth, sigma)
            kh = int(smooth / 2)                                   # the kernel receives int(smoothSize/2) (FB:213, 567)
            RA = polynomial_expansion(imresize_bilinear(gaussian_blur(im1, gk[:kh + 1]), width, height), gh, xgh, xxgh, ig)
            RB = polynomial_expansion(imresize_bilinear(gaussian_blur(im2, gk[:kh + 1]), width, height), gh, xgh, xxgh, ig)
            M = update_matrices(cur_x, cur_y, RA, RB)
            wk = blur_kernel_half(self.windowSize, self.windowSize / 2 * 0.3) if self.useGaussianFilter else None
            wh = int(self.windowSize / 2)
            for i in range(self.numIters):                         # FB:589-593
                M = _blur5(M, wk[:wh + 1], None) if self.useGaussianFilter else _blur5(M, None, wh)
                cur_x, cur_y = update_flow(M)
                if i < self.numIters - 1:
                    M = update_matrices(cur_x, cur_y, RA, RB)
            prev = (cur_x, cur_y)
        return cur_x, cur_y, "Unknown"
