"""CPU oracle for the HS + Liu-Shen pyramidal optical-flow path.  TEST INFRASTRUCTURE ONLY.

This is a from-scratch numpy restatement of the arithmetic of the reference
(alexlib/OpticalFlow-RI, path /root/reference) for ONE hot path: the
`genericPyramidalOpticalFlow` driver with the Horn-Schunck and Liu-Shen adapters and the
per-level stages.  It exists so that the CUDA path can be checked against something that
runs on a box without /root/reference.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product
(`opticalflow_ri_b200`) never imports it and has no CPU fallback.

Parity status: the reference ships NO tests and NO golden vectors of its own, so
the oracle is pinned against outputs of the unmodified reference run in the build container:
`oracle/make_golden.py` imports /root/reference/src and writes `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function below against those vectors
(bit-exact for gaussian / resize / warp / derivatives / Liu-Shen coefficients; <=1e-6 for the
spline and the iterative solvers).

The third-party arithmetic the reference delegates to (scipy.ndimage.convolve, Pillow
BICUBIC resize, FITPACK RectBivariateSpline, numba f32 loops) is restated here from its
published algorithm -- nothing below calls scipy, Pillow or numba.

Citations `GPOF/HS/LS/GF/GKBE:line` are into /root/reference/src/
GenericPyramidalOpticalFlow.py / HornSchunck.py / PhysicsBasedOpticalFlowLiuShen.py /
gaussian_filter.py / GaussianKernelBitExact.py.
"""
from __future__ import annotations

import math
from decimal import Decimal

import numpy as np

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------------------
# Gaussian pre-filter  (GF:24-94)
# --------------------------------------------------------------------------------------
def prepare_gaussian_kernel(sigma: float, ksize: int) -> np.ndarray:
    """GF:47-52.  f64 sampled Gaussian stored to f32, normalised by its f32 sum."""
    k = np.zeros(ksize, dtype=F32)
    xs = np.arange(-ksize / 2, ksize / 2, 1, dtype=int)  # [-h..h] for odd ksize
    k[:] = 1.0 / np.sqrt(2.0 * np.pi * sigma ** 2) * np.exp(-xs ** 2 / (2.0 * sigma ** 2))
    k /= np.sum(k)
    return k


def _gauss_pad_index(n: int, h: int) -> np.ndarray:
    """Source index of every sample of the padded line P (length n+2h), GF:63-66 / 75-78.

    left/top: P[h-1-j] = a[j] (mirror incl. edge); right/bottom: P[n+2h-1-j] = a[n-1-j], i.e. a
    forward copy of the last h samples (NOT a mirror -- reference quirk)."""
    idx = np.empty(n + 2 * h, dtype=np.int64)
    idx[h:h + n] = np.arange(n)
    for j in range(h):
        idx[h - 1 - j] = j
        idx[n + 2 * h - 1 - j] = n - 1 - j
    return idx


def _gauss_lines(a: np.ndarray, k: np.ndarray) -> np.ndarray:
    """Filter along the LAST axis.  out[x] = (((0 + P[x+2h]k0) + P[x+2h-1]k1) + ...) with separate
    f32 multiply and f32 add (GF:37-40: `result[i] += otherVec[i-j]*minVec[j]`, j ascending)."""
    n = a.shape[-1]
    K = k.shape[0]
    h = K // 2
    P = a[..., _gauss_pad_index(n, h)]
    acc = np.zeros(a.shape, dtype=F32)
    for j in range(K):
        lo = 2 * h - j
        acc = (acc + (P[..., lo:lo + n] * k[j]).astype(F32)).astype(F32)
    return acc


def gaussian_filter_kernel(img: np.ndarray, k: np.ndarray) -> np.ndarray:
    """GF:54-85 convolveSeparableFilter: rows first, then columns of the row-filtered image."""
    a = np.ascontiguousarray(img, dtype=F32)
    a = _gauss_lines(a, k)
    a = _gauss_lines(np.ascontiguousarray(a.T), k).T
    return np.ascontiguousarray(a)


def gaussian_filter_px(img: np.ndarray, sigma: float, ksize: int) -> np.ndarray:
    """GF:92-94 gaussian_filterPx (returns a new array; the reference works in place on a copy)."""
    return gaussian_filter_kernel(img, prepare_gaussian_kernel(sigma, ksize))


def gaussian_filter_truncate(img: np.ndarray, sigma: float, truncate: float) -> np.ndarray:
    """GF:87-90 gaussian_filter(image, sigma, truncate)."""
    ksize = 2 * int(truncate * sigma + 0.5) + 1
    return gaussian_filter_px(img, sigma, ksize)


# --------------------------------------------------------------------------------------
# Down-sampling: Pillow Image.resize(BICUBIC) on mode 'F'   (GPOF:67-68, third-party Resample.c)
# --------------------------------------------------------------------------------------
def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    if x < 2.0:
        return (((x - 5.0) * x + 8.0) * x - 4.0) * a
    return 0.0


def _bilinear(x: float) -> float:
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def resize_taps(in_size: int, out_size: int, kind: str = "bicubic"):
    """Per-output tap window [xmin, xmin+cnt) and normalised f64 weights (Pillow precompute_coeffs); kind = "bicubic"
    (support 2; the driver's down-sampling, GPOF:67-68) or "bilinear" (support 1; Farneback_PyCL.py:61-62)."""
    filt, fsup = (_bicubic, 2.0) if kind == "bicubic" else (_bilinear, 1.0)
    scale = in_size / out_size
    fs = max(scale, 1.0)
    sup = fsup * fs
    kmax = int(math.ceil(sup)) * 2 + 1
    xmin = np.zeros(out_size, dtype=np.int32)
    cnt = np.zeros(out_size, dtype=np.int32)
    w = np.zeros((out_size, kmax), dtype=F64)
    ss = 1.0 / fs
    for i in range(out_size):
        c = (i + 0.5) * scale
        lo = int(c - sup + 0.5)
        if lo < 0:
            lo = 0
        hi = int(c + sup + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        tot = 0.0
        for x in range(n):
            v = filt((x + lo - c + 0.5) * ss)
            w[i, x] = v
            tot += v
        for x in range(n):
            if tot != 0.0:
                w[i, x] /= tot
        xmin[i] = lo
        cnt[i] = n
    return xmin, cnt, w


def _resample_last_axis(a: np.ndarray, out_size: int, kind: str = "bicubic") -> np.ndarray:
    n = a.shape[-1]
    xmin, cnt, w = resize_taps(n, out_size, kind)
    out = np.empty(a.shape[:-1] + (out_size,), dtype=F32)
    a64 = a.astype(F64)
    for i in range(out_size):
        ss = np.zeros(a.shape[:-1], dtype=F64)
        for x in range(cnt[i]):
            ss = ss + a64[..., xmin[i] + x] * w[i, x]
        out[..., i] = ss.astype(F32)
    return out


def imresize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """GPOF:67-68: horizontal pass (f32 intermediate) then vertical pass; f64 accumulate ascending."""
    a = np.ascontiguousarray(img, dtype=F32)
    H, W = a.shape
    if out_w != W:
        a = _resample_last_axis(a, out_w)
    if out_h != H:
        a = np.ascontiguousarray(_resample_last_axis(np.ascontiguousarray(a.T), out_h).T)
    return np.ascontiguousarray(a)


def imresize_filter(img: np.ndarray, out_w: int, out_h: int, kind: str) -> np.ndarray:
    """Image.resize with another Pillow filter (same two-pass scheme); an unchanged size returns a copy, as Pillow does."""
    a = np.ascontiguousarray(img, dtype=F32)
    H, W = a.shape
    if out_w != W:
        a = _resample_last_axis(a, out_w, kind)
    if out_h != H:
        a = np.ascontiguousarray(_resample_last_axis(np.ascontiguousarray(a.T), out_h, kind).T)
    return np.ascontiguousarray(a).copy()


def level_size(n: int, scale: float) -> int:
    """GPOF:338-343: np.int32(np.round(n*scale)), round half to even."""
    return int(np.int32(np.round(n * scale)))


# --------------------------------------------------------------------------------------
# Flow up-sampling: RectBivariateSpline(kx=ky=3, s=0) == separable not-a-knot cubic spline
# (GPOF:152-172; third-party FITPACK regrid/bispev)
# --------------------------------------------------------------------------------------
def _notaknot_second_derivs(y: np.ndarray) -> np.ndarray:
    """Second derivatives M (unit knot spacing) of the not-a-knot cubic through y along axis 0.

    Interior: M[i-1] + 4 M[i] + M[i+1] = 6 (y[i-1] - 2 y[i] + y[i+1]);
    ends: M0 - 2 M1 + M2 = 0 and M[n-3] - 2 M[n-2] + M[n-1] = 0.
    Eliminating M0 and M[n-1] gives a tridiagonal system in M1..M[n-2] solved by the Thomas
    algorithm in f64."""
    n = y.shape[0]
    if n < 4:
        raise ValueError("cubic spline needs at least 4 samples per axis")
    y = y.astype(F64)
    m = n - 2
    rhs = 6.0 * (y[:-2] - 2.0 * y[1:-1] + y[2:])          # rows 1..n-2
    lo = np.ones(m, dtype=F64)
    di = np.full(m, 4.0, dtype=F64)
    up = np.ones(m, dtype=F64)
    # row 1: M0 = 2 M1 - M2  ->  6 M1 + 0 M2 = rhs
    di[0] = 6.0
    up[0] = 0.0
    # row n-2: M[n-1] = 2 M[n-2] - M[n-3]  ->  0 M[n-3] + 6 M[n-2] = rhs
    di[m - 1] = 6.0
    lo[m - 1] = 0.0
    cp = np.zeros(m, dtype=F64)
    dp = np.zeros((m,) + y.shape[1:], dtype=F64)
    cp[0] = up[0] / di[0]
    dp[0] = rhs[0] / di[0]
    for i in range(1, m):
        den = di[i] - lo[i] * cp[i - 1]
        cp[i] = up[i] / den
        dp[i] = (rhs[i] - lo[i] * dp[i - 1]) / den
    M = np.zeros_like(y)
    M[m] = dp[m - 1]
    for i in range(m - 2, -1, -1):
        M[i + 1] = dp[i] - cp[i] * M[i + 2]
    M[0] = 2.0 * M[1] - M[2]
    M[n - 1] = 2.0 * M[n - 2] - M[n - 3]
    return M


def _spline_axis0(y: np.ndarray, out_n: int) -> np.ndarray:
    """Evaluate the not-a-knot spline through y (axis 0, n samples at i/n) at k/out_n, clamped
    to the last sample (FITPACK clamps the argument to the knot range)."""
    n = y.shape[0]
    y = y.astype(F64)
    M = _notaknot_second_derivs(y)
    k = np.arange(out_n, dtype=np.int64)
    num = k * n
    i = num // out_n
    s = (num - i * out_n).astype(F64) / float(out_n)
    over = i >= n - 1
    i = np.where(over, n - 2, i)
    s = np.where(over, 1.0, s)
    sh = (out_n,) + (1,) * (y.ndim - 1)
    s = s.reshape(sh)
    t = 1.0 - s
    Mi, Mj, yi, yj = M[i], M[i + 1], y[i], y[i + 1]
    return Mi * (t * t * t) / 6.0 + Mj * (s * s * s) / 6.0 + (yi - Mi / 6.0) * t + (yj - Mj / 6.0) * s


def spline_upsample(a: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """GPOF:155-162: f32(RectBivariateSpline(i/h, j/w, a)(k/H, l/W)).  Returns f32 H x W."""
    t = _spline_axis0(np.asarray(a, dtype=F64), out_h)          # H x w
    t = _spline_axis0(np.ascontiguousarray(t.T), out_w).T       # H x W
    return np.ascontiguousarray(t.astype(F32))


# --------------------------------------------------------------------------------------
# Bilinear warp (GPOF:70-116) and level transition (GPOF:118-235, bilinear branch)
# --------------------------------------------------------------------------------------
def bilinear_warp(img: np.ndarray, cy: np.ndarray, cx: np.ndarray) -> np.ndarray:
    """GPOF:70-116.  cy/cx are f32 coordinate planes; weights are f64; clamp AFTER the fraction."""
    img = np.asarray(img, dtype=F32)
    H, W = img.shape
    cy = np.asarray(cy, dtype=F32)
    cx = np.asarray(cx, dtype=F32)
    iy = np.rint(cy).astype(np.int32)
    ix = np.rint(cx).astype(np.int32)
    dy = cy.astype(F64) - iy
    dx = cx.astype(F64) - ix
    ny = np.where(dy < 0, iy - 1, iy + 1)
    nx = np.where(dx < 0, ix - 1, ix + 1)
    dy = np.abs(dy)
    dx = np.abs(dx)
    iy = np.clip(iy, 0, H - 1)
    ny = np.clip(ny, 0, H - 1)
    ix = np.clip(ix, 0, W - 1)
    nx = np.clip(nx, 0, W - 1)
    out = ((1 - dy) * (1 - dx) * img[iy, ix] + (1 - dy) * dx * img[iy, nx]
           + dy * (1 - dx) * img[ny, ix] + dy * dx * img[ny, nx])
    return out.astype(F32)


def warp_pair(im1: np.ndarray, im2: np.ndarray, us: np.ndarray, vs: np.ndarray):
    """GPOF:186-201: im1 sampled at (y - v/2, x - u/2), im2 at (y + v/2, x + u/2)."""
    H, W = im1.shape
    ys = np.arange(H, dtype=np.int32)[:, None]
    xs = np.arange(W, dtype=np.int32)[None, :]
    hv = (vs / 2.0).astype(F32)      # f32 array / python float stays f32
    hu = (us / 2.0).astype(F32)
    w1 = bilinear_warp(im1, F32(ys - hv.astype(F64)), F32(xs - hu.astype(F64)))
    w2 = bilinear_warp(im2, F32(ys + hv.astype(F64)), F32(xs + hu.astype(F64)))
    return w1, w2


def upsample_flow(Uacc, Vacc, H, W, scale: bool):
    """GPOF:152-172: spline up-sample (skipped when sizes match) then optional *= f32(W/w), f32(H/h)."""
    h, w = Uacc.shape
    if (h, w) != (H, W):
        us = spline_upsample(Uacc, H, W)
        vs = spline_upsample(Vacc, H, W)
    else:
        us = np.array(Uacc, dtype=F32, copy=True)
        vs = np.array(Vacc, dtype=F32, copy=True)
    if scale:
        us = us * F32(F32(W) / F32(w))
        vs = vs * F32(F32(H) / F32(h))
    return us.astype(F32), vs.astype(F32)


# --------------------------------------------------------------------------------------
# scipy.ndimage.convolve restated (f64 accumulate in kernel memory order, f32 store)
# --------------------------------------------------------------------------------------
def liu_shen_warp(im1: np.ndarray, us: np.ndarray, vs: np.ndarray) -> np.ndarray:
    """The biLinear=False branch of updateNextPyramidalLevel (GPOF:190-196, 204-221): integer-shift SCATTER of frame 1
    along the rounded flow, then the optical-flow equation with the Gaussian-smoothed sub-pixel parts.

    Scatter: `im1[vsSwap, usSwap] = im1[ysMesh, xsMesh]` -- the right-hand side is a copy, sources are visited in
    row-major order and the LAST writer of a target wins; negative targets wrap around (numpy indexing), targets
    >= size (or < -size) raise IndexError.  Only frame 1 is warped.  Returns the new frame (the reference mutates its
    argument in place)."""
    im1 = np.ascontiguousarray(im1, dtype=F32)
    H, W = im1.shape
    fu = np.floor(us.astype(F32) + F32(0.5))                     # usNew + 0.5 stays float32 (NEP 50 weak scalar)
    fv = np.floor(vs.astype(F32) + F32(0.5))
    xs, ys = np.meshgrid(np.arange(W, dtype=np.int32), np.arange(H, dtype=np.int32))
    tx = (xs.astype(F64) + fu.astype(F64)).astype(np.int64)      # int32 + float32 -> float64 -> np.int32()
    ty = (ys.astype(F64) + fv.astype(F64)).astype(np.int64)
    if (tx >= W).any() or (tx < -W).any() or (ty >= H).any() or (ty < -H).any():
        raise IndexError("Liu-Shen warp: scatter target outside the frame")
    tx = np.where(tx < 0, tx + W, tx)
    ty = np.where(ty < 0, ty + H, ty)
    winner = np.full(H * W, -1, dtype=np.int64)
    np.maximum.at(winner, (ty * W + tx).ravel(), np.arange(H * W, dtype=np.int64))     # last writer in row-major order
    flat = im1.ravel()
    out = np.where(winner >= 0, flat[np.maximum(winner, 0)], flat).reshape(H, W).astype(F32)
    dU = (us - fu).astype(F32)
    dV = (vs - fv).astype(F32)
    mask_size = 3                                                # GPOF:210-212: sigma = 0.6*3, truncate = 4.0/0.6*3 -> 73 taps
    dU = gaussian_filter_truncate(dU, 0.6 * mask_size, 4.0 / 0.6 * mask_size)
    dV = gaussian_filter_truncate(dV, 0.6 * mask_size, 4.0 / 0.6 * mask_size)
    tdx = (out[0:-1, 1:] * dU[0:-1, 1:] - out[0:-1, 0:-1] * dU[0:-1, 0:-1]).astype(F32)
    tdy = (out[1:, 0:-1] * dV[1:, 0:-1] - out[0:-1, 0:-1] * dV[0:-1, 0:-1]).astype(F32)
    res = out.copy()
    res[0:-1, 0:-1] = (out[0:-1, 0:-1] - (tdx + tdy).astype(F32)).astype(F32)
    return res


def _pad(a: np.ndarray, mode: str) -> np.ndarray:
    if mode == "mirror":      # scipy 'mirror' == numpy 'reflect' (edge not repeated)
        return np.pad(a, 1, mode="reflect")
    if mode == "nearest":
        return np.pad(a, 1, mode="edge")
    if mode == "constant":
        return np.pad(a, 1, mode="constant")
    raise ValueError(mode)


def _correlate3(a: np.ndarray, w, mode: str) -> np.ndarray:
    """out[i,j] = f32( sum_{r,c row-major, w!=0} w[r][c] * a[i+r-1, j+c-1] ), f64 accumulate from 0."""
    H, W = a.shape
    p = _pad(a.astype(F64), mode)
    acc = np.zeros((H, W), dtype=F64)
    for r in range(3):
        for c in range(3):
            wv = float(w[r][c])
            if wv != 0.0:
                acc = acc + p[r:r + H, c:c + W] * wv
    return acc.astype(F32)


# --------------------------------------------------------------------------------------
# Horn-Schunck  (HS:52-127)
# --------------------------------------------------------------------------------------
def _mirror_next(a: np.ndarray) -> np.ndarray:
    """Pad one sample after the end of both axes with scipy 'mirror' (index N -> N-2)."""
    return np.pad(a, ((0, 1), (0, 1)), mode="reflect")


def hs_derivatives(frame1: np.ndarray, frame2: np.ndarray):
    """HS:107-127 as reached through HS:84 with the argument swap of HS:37/73.

    frame1/frame2 are the images as given to `compute(im1, im2, ...)`.  2x2 stencils over
    (i..i+1, j..j+1), each convolve f64-accumulated and stored f32, then combined in f32:
      fx = gx(frame2) + gx(frame1),  gx(I) = (I00 - I01 + I10 - I11)/4
      fy = gy(frame2) + gy(frame1),  gy(I) = (I00 + I01 - I10 - I11)/4
      ft = box(frame1) - box(frame2), box = mean of the 2x2 block."""
    A = _mirror_next(np.asarray(frame1, dtype=F32).astype(F64))
    B = _mirror_next(np.asarray(frame2, dtype=F32).astype(F64))

    def taps(P):
        return P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]

    def gx(P):
        i00, i01, i10, i11 = taps(P)
        # scipy visits the flipped 2x2 kernel row-major: (i,j), (i,j+1), (i+1,j), (i+1,j+1)
        return ((((0.0 + i00 * 0.25) + i01 * -0.25) + i10 * 0.25) + i11 * -0.25).astype(F32)

    def gy(P):
        i00, i01, i10, i11 = taps(P)
        return ((((0.0 + i00 * 0.25) + i01 * 0.25) + i10 * -0.25) + i11 * -0.25).astype(F32)

    def box(P, s):
        i00, i01, i10, i11 = taps(P)
        return ((((0.0 + i00 * s) + i01 * s) + i10 * s) + i11 * s).astype(F32)

    fx = gx(B) + gx(A)
    fy = gy(B) + gy(A)
    ft = box(A, 0.25) + box(B, -0.25)
    return fx.astype(F32), fy.astype(F32), ft.astype(F32)


_HS_KERNEL = np.array([[1 / 12, 1 / 6, 1 / 12], [1 / 6, 0, 1 / 6], [1 / 12, 1 / 6, 1 / 12]], dtype=F32)


def hs_iterate(U, V, fx, fy, ft, alpha, niter: int):
    """HS:62-71 + HS:52-59: exactly `niter` Jacobi sweeps; stencil f64-acc -> f32, update in f32."""
    U = np.asarray(U, dtype=F32)
    V = np.asarray(V, dtype=F32)
    a2 = F32(F32(alpha) * F32(alpha))
    den = (a2 + fx * fx + fy * fy).astype(F32)
    for _ in range(int(niter)):
        ua = _correlate3(U, _HS_KERNEL, "mirror")
        va = _correlate3(V, _HS_KERNEL, "mirror")
        der = ((fx * ua + fy * va + ft) / den).astype(F32)
        U = (ua - fx * der).astype(F32)
        V = (va - fy * der).astype(F32)
    return U, V


def hs_error(Unew, Vnew, U0, V0) -> float:
    """HS:100: (||Unew-U0||_F + ||Vnew-V0||_F) / (H*W)."""
    H, W = Unew.shape
    du = (Unew - np.asarray(U0, dtype=F32)).astype(F64)
    dv = (Vnew - np.asarray(V0, dtype=F32)).astype(F64)
    return float((np.sqrt(np.sum(du * du)) + np.sqrt(np.sum(dv * dv))) / (H * W))


def hs_compute(im1, im2, alpha, niter, U, V):
    """HSOpticalFlowAlgoAdapter.compute for one alpha (HS:35-37 -> HS:73-105)."""
    fx, fy, ft = hs_derivatives(im1, im2)
    Un, Vn = hs_iterate(U, V, fx, fy, ft, alpha, niter)
    return Un, Vn, hs_error(Un, Vn, U, V)


# --------------------------------------------------------------------------------------
# Liu-Shen  (LS:33-158)
# --------------------------------------------------------------------------------------
_D_ROW = [[0, -0.5, 0], [0, 0, 0], [0, 0.5, 0]]       # (x[i+1,j] - x[i-1,j]) / 2
_D_COL = [[0, 0, 0], [-0.5, 0, 0.5], [0, 0, 0]]       # (x[i,j+1] - x[i,j-1]) / 2
_F_ROW = [[0, 1, 0], [0, 0, 0], [0, 1, 0]]            # x[i-1,j] + x[i+1,j]
_F_COL = [[0, 0, 0], [1, 0, 1], [0, 0, 0]]
_MIX = [[0.25, 0, -0.25], [0, 0, 0], [-0.25, 0, 0.25]]
_D2_ROW = [[0, 1, 0], [0, -2, 0], [0, 1, 0]]
_D2_COL = [[0, 0, 0], [1, -2, 1], [0, 0, 0]]
_H8 = [[1, 1, 1], [1, 0, 1], [1, 1, 1]]


def ls_coefficients(im1, im2, h):
    """LS:96-97, 124-128 and generate_invmatrix LS:47-73.  All products/sums in f32."""
    im1 = np.asarray(im1, dtype=F32)
    im2 = np.asarray(im2, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        i1 = (im1 / np.max(im1)).astype(F32)
        i2 = (im2 / np.max(im2)).astype(F32)
        IIx = i1 * _correlate3(i1, _D_ROW, "nearest")
        IIy = i1 * _correlate3(i1, _D_COL, "nearest")
        II = i1 * i1
        dI = (i2 - i1).astype(F32)
        Ixt = i1 * _correlate3(dI, _D_ROW, "nearest")
        Iyt = i1 * _correlate3(dI, _D_COL, "nearest")
        hh = F32(h)
        cm = _correlate3(np.ones(i1.shape, dtype=F32), _H8, "constant")
        A11 = i1 * (_correlate3(i1, _D2_ROW, "nearest") - 2 * i1) - hh * cm
        A22 = i1 * (_correlate3(i1, _D2_COL, "nearest") - 2 * i1) - hh * cm
        A12 = i1 * _correlate3(i1, _MIX, "nearest")
        det = A11 * A22 - A12 * A12
        B11 = A22 / det
        B12 = -A12 / det
        B22 = A11 / det
    return tuple(np.asarray(x, dtype=F32) for x in (IIx, IIy, II, Ixt, Iyt, B11, B12, B22))


def ls_iterate(u, v, coef, h, maxnum=60, tol=1e-8):
    """LS:141-156 + helper LS:75-80.  `u` is the ROW (V) component, `v` the COLUMN (U) component."""
    IIx, IIy, II, Ixt, Iyt, B11, B12, B22 = coef
    u = np.asarray(u, dtype=F32)
    v = np.asarray(v, dtype=F32)
    r, c = u.shape
    total_error = 1e8
    err = 0.0
    k = 0
    hf = F32(h)
    while total_error > tol and k < maxnum:
        bu = (2 * IIx * _correlate3(u, _D_ROW, "nearest") + IIx * _correlate3(v, _D_COL, "nearest")
              + IIy * _correlate3(v, _D_ROW, "nearest") + II * _correlate3(u, _F_ROW, "nearest")
              + II * _correlate3(v, _MIX, "nearest") + hf * _correlate3(u, _H8, "constant") + Ixt)
        bv = (IIy * _correlate3(u, _D_ROW, "nearest") + IIx * _correlate3(u, _D_COL, "nearest")
              + 2 * IIy * _correlate3(v, _D_COL, "nearest") + II * _correlate3(u, _MIX, "nearest")
              + II * _correlate3(v, _F_COL, "nearest") + hf * _correlate3(v, _H8, "constant") + Iyt)
        unew = (-(B11 * bu + B12 * bv)).astype(F32)
        vnew = (-(B12 * bu + B22 * bv)).astype(F32)
        du = (unew - u).astype(F64)
        dv = (vnew - v).astype(F64)
        total_error = float((F32(np.sqrt(np.sum(du * du))) + F32(np.sqrt(np.sum(dv * dv)))) / (r * c))
        u, v = unew, vnew
        err = total_error
        k += 1
    return u, v, err, k


def ls_compute(im1, im2, h, U, V, maxnum=60, tol=1e-8):
    """LiuShenOpticalFlowAlgoAdapter.compute (LS:37-39): U/V swapped on the way in and out."""
    coef = ls_coefficients(im1, im2, h)
    with np.errstate(all="ignore"):
        u, v, err, k = ls_iterate(V, U, coef, h, maxnum, tol)
    return v, u, err, k          # (U, V, error, iterations)


# --------------------------------------------------------------------------------------
# Driver  (GPOF:238-416), bilinear-warp / no-warp branches, native adapters only
# --------------------------------------------------------------------------------------
class HSParams:
    """Oracle-side stand-in for HSOpticalFlowAlgoAdapter (HS:29-50)."""
    name = "HS"

    def __init__(self, alphas, niter, provide_defaults=True):
        self.alphas = list(alphas)
        self.niter = int(niter)
        self.provide_defaults = provide_defaults

    def defaults(self):
        return {"warping": True, "biLinear": True, "scaling": True} if self.provide_defaults else None

    def compute(self, im1, im2, U, V):
        alpha = self.alphas.pop()           # HS:36 -- LAST alpha first
        return hs_compute(im1, im2, alpha, self.niter, U, V)


class LSParams:
    """Oracle-side stand-in for LiuShenOpticalFlowAlgoAdapter (LS:33-45)."""
    name = "LS"

    def __init__(self, h, maxnum=60, tol=1e-8):
        self.h = h
        self.maxnum = maxnum
        self.tol = tol

    def defaults(self):
        return None

    def compute(self, im1, im2, U, V):
        U, V, err, _ = ls_compute(im1, im2, self.h, U, V, self.maxnum, self.tol)
        return U, V, err


def pyramidal_flow(im1, im2, FILTER, main, pyramidalLevels=1, kLevels=1, FILTER_OPT=None, optional=None,
                   warping=True, biLinear=True, pyramidalIntermediateScaling=True, pyramidalScaling=False,
                   trace=None):
    """genericPyramidalOpticalFlow (GPOF:238-416).  `trace`, if a list, receives per-level dicts."""
    im1 = np.ascontiguousarray(im1, dtype=F32)
    im2 = np.ascontiguousarray(im2, dtype=F32)
    d = main.defaults()
    if d is not None:                                            # GPOF:304-327
        warping = d.get("warping", warping) if d.get("warping") is not None else warping
        biLinear = d.get("biLinear", biLinear) if d.get("biLinear") is not None else biLinear
        if d.get("intermediateScaling") is not None:
            pyramidalIntermediateScaling = d["intermediateScaling"]
        if d.get("scaling") is not None:
            pyramidalScaling = d["scaling"]
    if optional is not None and FILTER_OPT is None:
        raise TypeError("'>' not supported between instances of 'NoneType' and 'float'")   # GPOF:380
    scale = 1.0 / (2.0 ** (pyramidalLevels - 1))
    U = V = Uacc = Vacc = None
    work1 = work2 = None
    H0, W0 = im1.shape
    for level in range(1, pyramidalLevels + 1):
        local_scaling = pyramidalScaling if level == pyramidalLevels else pyramidalIntermediateScaling
        if scale < 1.0 and level != pyramidalLevels:
            w = level_size(W0, scale)
            h = level_size(H0, scale)
            n1 = imresize_bicubic(im1, w, h)
            n2 = imresize_bicubic(im2, w, h)
        elif scale > 1.0:
            raise Exception("Invalid scale level: " + str(scale))
        else:
            n1, n2 = im1, im2
        Hl, Wl = n1.shape
        if level > 1:
            us, vs = upsample_flow(Uacc, Vacc, Hl, Wl, local_scaling)
            if warping and not biLinear:                          # GPOF:204-221: frame 1 is warped IN PLACE, frame 2 is not
                n1 = liu_shen_warp(n1, us, vs)
                if level == pyramidalLevels:
                    im1 = n1                                      # ... which at the last level is the caller's frame
                w1, w2 = n1, n2
                U = np.zeros((Hl, Wl), dtype=F32)
                V = np.zeros((Hl, Wl), dtype=F32)
                Uacc, Vacc = us, vs
            elif warping:
                w1, w2 = warp_pair(n1, n2, us, vs)
                U = np.zeros((Hl, Wl), dtype=F32)
                V = np.zeros((Hl, Wl), dtype=F32)
                Uacc, Vacc = us, vs
            else:                                                 # GPOF:228-232
                w1, w2 = n1, n2
                U, V = us, vs
                Uacc = np.zeros((Hl, Wl), dtype=F32)
                Vacc = np.zeros((Hl, Wl), dtype=F32)
        else:
            w1, w2 = n1, n2
            U = np.zeros((Hl, Wl), dtype=F32)
            V = np.zeros((Hl, Wl), dtype=F32)
            Uacc = np.zeros((Hl, Wl), dtype=F32)
            Vacc = np.zeros((Hl, Wl), dtype=F32)
        if FILTER > 1e-3:
            work1 = gaussian_filter_px(w1, FILTER, 3)
            work2 = gaussian_filter_px(w2, FILTER, 3)
        else:
            work1, work2 = w1.copy(), w2
        if optional is not None and FILTER_OPT > 1e-3:
            opt1 = gaussian_filter_px(n1, FILTER_OPT, 5)
            opt2 = gaussian_filter_px(n2, FILTER_OPT, 5)
        elif optional is not None:
            opt1, opt2 = n1.copy(), n2
        for k in range(kLevels):
            if k > 0:                                             # GPOF:392-404
                if warping:
                    us, vs = upsample_flow(Uacc, Vacc, Hl, Wl, False)
                    if biLinear:
                        w1, w2 = warp_pair(n1, n2, us, vs)
                    else:
                        w1, w2 = liu_shen_warp(n1, us, vs), n2    # warps a COPY this time (GPOF:394)
                    U = np.zeros((Hl, Wl), dtype=F32)
                    V = np.zeros((Hl, Wl), dtype=F32)
                    Uacc, Vacc = us, vs
                    if FILTER > 1:
                        work1 = gaussian_filter_px(w1, FILTER, 3)
                        work2 = gaussian_filter_px(w2, FILTER, 3)
                    else:
                        work1, work2 = w1.copy(), w2
                else:
                    U, V = Uacc.copy(), Vacc.copy()
                    Uacc = np.zeros((Hl, Wl), dtype=F32)
                    Vacc = np.zeros((Hl, Wl), dtype=F32)
            U, V, err_main = main.compute(work1, work2, U, V)
            err_opt = None
            if optional is not None:
                U, V, err_opt = optional.compute(opt1.copy(), opt2.copy(), U, V)
            Uacc = (Uacc + U).astype(F32)
            Vacc = (Vacc + V).astype(F32)
            if trace is not None:
                trace.append(dict(level=level, k=k, work1=work1, work2=work2, U=U, V=V, Uacc=Uacc.copy(),
                                  Vacc=Vacc.copy(), err_main=err_main, err_opt=err_opt))
        scale *= 2
    return Uacc, Vacc


# --------------------------------------------------------------------------------------
# getGaussianKernelBitExact  (GKBE:55-144) -- host-only coefficient generator
# --------------------------------------------------------------------------------------
_GKBE_FIXED = {
    1: [1.0],
    3: [0.25, 0.5, 0.25],
    5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
    7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125],
    9: [4 / 256, 13 / 256, 30 / 256, 51 / 256, 60 / 256, 51 / 256, 30 / 256, 13 / 256, 4 / 256],
}


def gaussian_kernel_bit_exact(n: int, sigma: float):
    """GKBE:55-144.  Decimal (28 digits) arithmetic; a POSITIVE sigma is ignored (GKBE:102-107)."""
    assert n > 0
    if sigma <= 0 and n in _GKBE_FIXED:
        return 1.0, np.array(_GKBE_FIXED[n], dtype=F64)
    one = Decimal("1.0")
    sx = Decimal(sigma) if sigma < 0 else Decimal(n) * Decimal("0.15") + Decimal("0.35")
    scale2x = Decimal("-0.125") / (sx * sx)
    n2 = int((n - 1) / 2)
    vals = []
    x = 1 - n
    tot = Decimal("0.0")
    for _ in range(n2):
        t = (Decimal(x * x) * scale2x).exp()
        vals.append(t)
        tot += t
        x += 2
    tot *= Decimal(2.0)
    tot += one
    if (n & 1) == 0:
        tot += one
    mul1 = one / tot
    res = [Decimal(0)] * n
    sum2 = Decimal("0.0")
    for i in range(n2):
        t = vals[i] * mul1
        res[i] = t
        res[n - 1 - i] = t
        sum2 += t
    sum2 *= Decimal(2.0)
    res[n2] = one * mul1
    sum2 += res[n2]
    if (n & 1) == 0:
        res[n2 + 1] = res[n2]
        sum2 += res[n2]
    return float(sum2), np.array([float(r) for r in res], dtype=F64)


# --------------------------------------------------------------------------------------
# Synthetic PIV pair with known Poiseuille truth -- inputs for bench / property tests
# --------------------------------------------------------------------------------------
def poiseuille_truth(H: int, W: int):
    y = np.arange(H, dtype=F64)
    u = -4.0 * (1.0 - ((y - (H - 1) / 2.0) / (H / 2.0)) ** 2)
    return np.repeat(u[:, None], W, axis=1).astype(F32), np.zeros((H, W), dtype=F32)


def synthetic_piv_pair(H: int, W: int, seed: int = 0):
    """8-bit particle images: 6 particles / 256 px, Gaussian blobs sigma 0.75 px, displaced by
    -/+ u(y)/2 in frame 0/1.  Returns two float32 H x W arrays (values 0..255)."""
    rng = np.random.default_rng(seed)
    n = (H * W * 6) // 256
    px = rng.uniform(-8.0, W + 8.0, n)
    py = rng.uniform(-8.0, H + 8.0, n)
    peak = 255.0 * np.exp(-0.5 * rng.standard_normal(n) ** 2)
    uy = -4.0 * (1.0 - ((py - (H - 1) / 2.0) / (H / 2.0)) ** 2)
    frames = []
    R = 4
    offs = np.arange(-R, R + 1)
    for sgn in (-0.5, 0.5):
        img = np.zeros((H, W), dtype=F64)
        x = px + sgn * uy
        x0 = np.rint(x).astype(np.int64)
        y0 = np.rint(py).astype(np.int64)
        for dy in offs:
            yy = y0 + dy
            wy = np.exp(-((yy - py) ** 2) / (2 * 0.75 ** 2))
            oky = (yy >= 0) & (yy < H)
            for dx in offs:
                xx = x0 + dx
                ok = oky & (xx >= 0) & (xx < W)
                val = peak * wy * np.exp(-((xx - x) ** 2) / (2 * 0.75 ** 2))
                np.add.at(img, (yy[ok], xx[ok]), val[ok])
        frames.append(np.clip(np.rint(img), 0, 255).astype(np.uint8).astype(F32))
    return frames[0], frames[1]


def epe_rmse(U, V, Ut, Vt) -> float:
    d = (U.astype(F64) - Ut) ** 2 + (V.astype(F64) - Vt) ** 2
    return float(np.sqrt(np.mean(d)))
