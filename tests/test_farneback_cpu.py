"""CPU tests of the Farneback adapter's host logic and of its oracle (oracle/ofri_farneback_oracle.py).

The reference's adapter computes in OpenCL kernels that cannot run in this image, so the oracle is pinned in two layers:
the coefficient tables against the REFERENCE's own generators (tests/golden/farneback_tables.npz, made by
`oracle/make_golden.py --farneback`), and the algorithm against OpenCV's CPU implementation that the reference's kernels
port (cv2.calcOpticalFlowFarneback), when cv2 is importable."""
import sys

import numpy as np
import pytest

import ofri_farneback_oracle as FBO
import ofri_oracle as O

TABLE_CASES = {"7_15": (7, 1.5), "5_11": (5, 1.1), "5_0": (5, 0.0), "7_12": (7, 1.2)}
BLUR_CASES = {"33": (33, 33 / 2 * 0.3), "13": (13, 13 / 2 * 0.3), "3_0": (3, 0.0), "3_05": (3, 0.5), "7_15": (7, 1.5),
              "17_35": (17, 3.5)}


@pytest.fixture(scope="module")
def tables():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "farneback_tables.npz"))


@pytest.fixture(scope="module")
def FB():
    import opticalflow_ri_b200 as ofri
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import Farneback_PyCL as m
    finally:
        sys.path.remove(ofri.SRC_DIR)
    return m


def piv_pair(seed, H, W, shift=(1.3, -0.7)):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    n = H * W // 30
    cx, cy, a = rng.uniform(0, W, n), rng.uniform(0, H, n), rng.uniform(60, 220, n)

    def render(dx, dy):
        im = np.zeros((H, W))
        for x0, y0, a0 in zip(cx + dx, cy + dy, a):
            x_lo, x_hi = max(int(x0) - 5, 0), min(int(x0) + 6, W)
            y_lo, y_hi = max(int(y0) - 5, 0), min(int(y0) + 6, H)
            if x_lo < x_hi and y_lo < y_hi:
                im[y_lo:y_hi, x_lo:x_hi] += a0 * np.exp(-((xx[y_lo:y_hi, x_lo:x_hi] - x0) ** 2 +
                                                          (yy[y_lo:y_hi, x_lo:x_hi] - y0) ** 2) / (2 * 1.6 ** 2))
        return np.clip(im, 0, 255).astype(np.float32)

    return render(0, 0), render(*shift)


def test_oracle_tables_match_reference(tables):
    for tag, (n, sigma) in TABLE_CASES.items():
        g, xg, xxg, i11, i03, i33, i55 = FBO.prepare_gaussian(n, sigma)
        assert np.array_equal(g[n:], tables["g_" + tag]) and np.array_equal(xg[n:], tables["xg_" + tag])
        assert np.array_equal(xxg[n:], tables["xxg_" + tag])
        assert np.array_equal(np.float64([i11, i03, i33, i55]), tables["igd_" + tag])
        assert np.array_equal(np.float64([i11, i03, i33, i55]).astype(np.float32), tables["ig_" + tag])
    for tag, (size, sigma) in BLUR_CASES.items():
        assert np.array_equal(FBO.blur_kernel_half(size, sigma)[:size // 2 + 1], tables["blur_" + tag])


def test_dropin_tables_match_reference(FB, tables):
    """The product's own table generators (src/Farneback_PyCL.py) against the reference's, and the packed C struct."""
    for tag, (n, sigma) in TABLE_CASES.items():
        a = FB.Farneback_PyCL(polyN=n, polySigma=sigma)
        g, xg, xxg, i11, i03, i33, i55 = a.FarnebackPrepareGaussian()
        assert np.array_equal(g[n:], tables["g_" + tag]) and np.array_equal(xg[n:], tables["xg_" + tag])
        assert np.array_equal(xxg[n:], tables["xxg_" + tag])
        assert np.array_equal(np.float64([i11, i03, i33, i55]), tables["igd_" + tag])
        p = a.native_params()
        assert np.array_equal(np.float32(list(p.g)[:n + 1]), tables["g_" + tag])
        assert np.array_equal(np.float32(list(p.ig)), tables["ig_" + tag])
    for tag, (size, sigma) in BLUR_CASES.items():
        assert np.array_equal(FB.Farneback_PyCL()._kernel_half(size, sigma), tables["blur_" + tag])
    a = FB.Farneback_PyCL(windowSize=13, pyramidalLevels=3)
    p = a.native_params()
    assert (p.window_size, p.n_iters, p.poly_n, p.use_gaussian, p.extra_levels) == (13, 5, 7, 1, 2)
    assert np.array_equal(np.float32(list(p.win_kernel)[:7]), tables["blur_13"])
    # per-level pre-blur: level 0 sigma 0 -> 3 taps; level 1 sigma 0.5 -> 3 taps; level 2 sigma 1.5 -> round(7.5) | 1 = 9 taps
    assert list(p.n_blur)[:3] == [1, 1, 4]
    assert np.array_equal(np.float32(list(p.blur_kernel[0])[:2]), tables["blur_3_0"])
    assert np.array_equal(np.float32(list(p.blur_kernel[1])[:2]), tables["blur_3_05"])
    assert np.array_equal(np.float32(list(p.blur_kernel[2])[:5]), FBO.blur_kernel_half(9, 1.5)[:5])


def test_dropin_api_mirrors_reference(FB):
    import inspect
    assert list(inspect.signature(FB.Farneback_PyCL.__init__).parameters)[1:] == [
        "windowSize", "Niters", "polyN", "polySigma", "useGaussian", "pyrScale", "pyramidalLevels", "platformID",
        "deviceID", "provideGenericPyramidalDefaults"]
    a = FB.Farneback_PyCL()
    assert a.getAlgoName() == "Farneback CL" and a.hasGenericPyramidalDefaults()
    assert a.getGenericPyramidalDefaults() == {"warping": False, "scaling": True}
    assert (a.windowSize, a.numIters, a.polyN, a.polySigma, a.pyrScale, a.pyramidalLevels) == (33, 5, 7, 1.5, 0.5, 0)
    with pytest.raises(Exception, match="odd"):
        FB.Farneback_PyCL(windowSize=32)
    with pytest.raises(AssertionError):
        FB.Farneback_PyCL(pyramidalLevels=0)


def test_oracle_blur_is_a_reflect101_separable_filter():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    im = rng.uniform(0, 255, (41, 53)).astype(np.float32)
    half = FBO.blur_kernel_half(7, 1.5)[:4]
    full = np.concatenate([half[:0:-1], half]).astype(np.float32)
    want = cv2.sepFilter2D(im, cv2.CV_32F, full, full, borderType=cv2.BORDER_REFLECT_101)
    got = FBO.gaussian_blur(im, half)
    assert np.abs(got - want).max() < 2e-4          # same taps; OpenCV sums in a different order


def test_oracle_bilinear_resample_is_pillow():
    from PIL import Image
    rng = np.random.default_rng(5)
    im = rng.uniform(-3, 3, (37, 45)).astype(np.float32)
    for (w, h) in ((23, 19), (45, 37), (90, 74), (51, 40)):
        want = np.array(Image.fromarray(im).resize((w, h), Image.BILINEAR))
        assert np.array_equal(FBO.imresize_bilinear(im, w, h), want)


@pytest.mark.parametrize("case", [(13, 6, 5, 1.1), (13, 6, 7, 1.5), (21, 8, 7, 1.5), (9, 6, 5, 1.2)])
def test_oracle_matches_opencv_at_the_fixed_point(case):
    """Same algorithm as OpenCV's CPU Farneback with the box window, compared where the two must agree: OpenCV's CPU code
    refreshes the G / h matrices row by row WHILE an iteration runs (so it is about one iteration ahead of its own OpenCL
    version, which the reference ports and the oracle follows: cv2 after 1 iteration = oracle after 2 to 7e-3 px), but both
    iterate the same map, so after 6-8 iterations they sit on the same fixed point.  Measured: max 7e-7 .. 2.6e-6 px.
    (The Gaussian-window variant differs by design: the reference's window comes from getGaussianKernelBitExact, which
    replaces a positive sigma by 0.15 n + 0.35.)"""
    cv2 = pytest.importorskip("cv2")
    win, iters, poly_n, poly_sigma = case
    a, b = piv_pair(11, 96, 112)
    z = np.zeros_like(a)
    fb = FBO.FBParams(windowSize=win, Niters=iters, polyN=poly_n, polySigma=poly_sigma, useGaussian=False, pyramidalLevels=1)
    U, V, _ = fb.compute(a, b, z, z)
    flow = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 1, win, iters, poly_n, poly_sigma, 0)
    m = win // 2 + 8
    d = np.abs(np.dstack([U, V]) - flow)[m:-m, m:-m]
    assert d.max() < 2e-5, d.max()
    assert abs(np.median(U[m:-m, m:-m]) - 1.3) < 0.15 and abs(np.median(V[m:-m, m:-m]) + 0.7) < 0.15


@pytest.mark.parametrize("case", [(13, 6, 5, 1.1), (21, 8, 7, 1.5)])
def test_oracle_gaussian_window_path_matches_opencv(case, monkeypatch):
    """The Gaussian-window code path (gaussianBlur5) against OpenCV's OPTFLOW_FARNEBACK_GAUSSIAN at the fixed point, with
    OpenCV's window kernel (sigma = 0.3 * (window / 2), float32 taps normalised once) substituted for the reference's:
    the reference builds its window with getGaussianKernelBitExact, whose values (pinned on the reference's own generator,
    test_oracle_tables_match_reference) are a different Gaussian, so only the PATH can be compared with OpenCV.
    Measured: max 1.3e-6 / 2.5e-6 px."""
    cv2 = pytest.importorskip("cv2")
    win, iters, poly_n, poly_sigma = case
    orig = FBO.blur_kernel_half

    def opencv_window(ksize, sigma):
        if ksize <= 3:                      # the level-0 pre-blur kernel: leave it alone
            return orig(ksize, sigma)
        m = ksize // 2
        s = m * 0.3
        k = np.exp(-np.arange(-m, m + 1, dtype=np.float64) ** 2 / (2 * s * s)).astype(np.float32)
        k = (k * np.float32(1.0 / float(np.sum(k, dtype=np.float64)))).astype(np.float32)
        return k[m:].copy()

    monkeypatch.setattr(FBO, "blur_kernel_half", opencv_window)
    a, b = piv_pair(11, 96, 112)
    z = np.zeros_like(a)
    fb = FBO.FBParams(windowSize=win, Niters=iters, polyN=poly_n, polySigma=poly_sigma, useGaussian=True, pyramidalLevels=1)
    U, V, _ = fb.compute(a, b, z, z)
    flow = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 1, win, iters, poly_n, poly_sigma, cv2.OPTFLOW_FARNEBACK_GAUSSIAN)
    m = win // 2 + 8
    d = np.abs(np.dstack([U, V]) - flow)[m:-m, m:-m]
    assert d.max() < 2e-5, d.max()


def test_oracle_internal_pyramid_recovers_the_shift():
    a, b = piv_pair(11, 96, 112)
    z = np.zeros_like(a)
    fb = FBO.FBParams(windowSize=13, Niters=3, polyN=5, polySigma=1.1, useGaussian=False, pyramidalLevels=2)
    U, V, _ = fb.compute(a, b, z, z)
    m = 24
    assert abs(np.median(U[m:-m, m:-m]) - 1.3) < 0.15 and abs(np.median(V[m:-m, m:-m]) + 0.7) < 0.15


def test_oracle_driver_accepts_farneback_adapter():
    """FB as main adapter of the generic driver (defaults warping False / scaling True), Liu-Shen as refinement."""
    a, b = piv_pair(2, 64, 72, shift=(0.8, 0.4))
    fb = FBO.FBParams(windowSize=13, Niters=2, polyN=5, polySigma=1.1)
    U, V = O.pyramidal_flow(a, b, 0.0, fb, 2, 1)[:2]
    assert U.shape == a.shape and np.isfinite(U).all() and np.isfinite(V).all()
    m = 16
    assert abs(np.median(U[m:-m, m:-m]) - 0.8) < 0.2 and abs(np.median(V[m:-m, m:-m]) - 0.4) < 0.2
