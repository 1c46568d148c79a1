"""CPU tests of the dense Lucas-Kanade adapter's host logic and of its oracle (oracle/ofri_lk_oracle.c).

The reference's adapter is one OpenCL kernel whose result depends on device-defined arithmetic (sampler filter
precision, mad, division), so there is nothing to pin bit for bit: the oracle fixes those three choices (see its header)
and is pinned on OpenCV's pyramidal Lucas-Kanade tracker run at level 0 on every pixel -- the algorithm the OpenCL kernel
was ported from -- and on the algebra of a translated image."""
import sys

import numpy as np
import pytest

import ofri_lk_oracle as LKO
from test_farneback_cpu import piv_pair


@pytest.fixture(scope="module")
def LK():
    import opticalflow_ri_b200 as ofri
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import denseLucasKanade_PyCL as m
    finally:
        sys.path.remove(ofri.SRC_DIR)
    return m


@pytest.mark.parametrize("half_window", [13, 7, 5, 15])      # 27 / 15 / 11 (the kernel's other weight rule) / 31 px windows
def test_oracle_matches_opencv_level0(half_window):
    cv2 = pytest.importorskip("cv2")
    a, b = piv_pair(3, 72, 80, shift=(1.3, -0.7))
    a8, b8 = np.clip(a, 0, 255).astype(np.uint8), np.clip(b, 0, 255).astype(np.uint8)
    ys, xs = np.mgrid[0:72, 0:80]
    pts = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32).reshape(-1, 1, 2)
    win = 2 * half_window + 1
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(a8, b8, pts, None, winSize=(win, win), maxLevel=0,
                                          criteria=(cv2.TERM_CRITERIA_COUNT | cv2.TERM_CRITERIA_EPS, 5, 0.01))
    want = (nxt - pts).reshape(72, 80, 2)
    z = np.zeros((72, 80), np.float32)
    u, v = LKO.lk_compute(a8.astype(np.float32), b8.astype(np.float32), z, z, 5, half_window)
    m = half_window + 4      # the window of a pixel closer than this to the border is clamped differently by the two
    d = np.abs(np.dstack([u, v]) - want)[m:-m, m:-m]
    # measured: max 1.0e-3 / 1.1e-3 / 2.1e-3 / 1.0e-3 px, median 2.7e-5 .. 3.1e-5 px
    assert d.max() < 5e-3 and np.median(d) < 1e-4, (d.max(), np.median(d))


def test_oracle_recovers_translation_and_keeps_flat_pixels():
    a, b = piv_pair(5, 64, 72, shift=(0.9, 0.5))
    z = np.zeros_like(a)
    u, v = LKO.lk_compute(a, b, z, z, 5, 13)
    m = 16
    assert abs(np.median(u[m:-m, m:-m]) - 0.9) < 0.05 and abs(np.median(v[m:-m, m:-m]) - 0.5) < 0.05
    # a flat frame has a singular structure tensor everywhere: the incoming flow is returned untouched (CL:478-484)
    flat = np.full((40, 48), 7.0, np.float32)
    u0 = np.random.default_rng(1).normal(0, 1, flat.shape).astype(np.float32)
    u, v = LKO.lk_compute(flat, flat, u0, -u0, 5, 13)
    assert np.array_equal(u, u0) and np.array_equal(v, -u0)
    # zero iterations: the flow goes through the kernel's (j + u - hw) + hw - j arithmetic only
    u, v = LKO.lk_compute(a, b, u0[:1, :1].repeat(64, 0).repeat(72, 1), z, 0, 13)
    assert np.abs(u - u0[0, 0]).max() < 1e-5 and np.abs(v).max() < 1e-5


def test_oracle_small_window_and_asymmetric_switches():
    a, b = piv_pair(6, 56, 60, shift=(0.6, -0.4))
    z = np.zeros_like(a)
    base = LKO.lk_compute(a, b, z, z, 5, 5)[0]              # 11 x 11 window: the kernel's WSX = WSY = 0 branch
    m = 12
    assert abs(np.median(base[m:-m, m:-m]) - 0.6) < 0.1
    for asym in ((0, 1, 0, 1), (1, 0, 0, 1)):
        u = LKO.lk_compute(a, b, z, z, 5, 13, asym)[0]
        ref = LKO.lk_compute(a, b, z, z, 5, 13)[0]
        assert not np.array_equal(u, ref) and abs(np.median(u[m:-m, m:-m]) - 0.6) < 0.1


def test_dropin_api_and_vorticity_switch(LK):
    import inspect
    assert list(inspect.signature(LK.denseLucasKanade_PyCl.__init__).parameters)[1:] == [
        "platformID", "deviceID", "Niter", "halfWindow", "provideGenericPyramidalDefaults", "enableVorticityEnhancement"]
    a = LK.denseLucasKanade_PyCl()
    assert a.getAlgoName() == "OpenCL Dense LK" and a.hasGenericPyramidalDefaults()
    assert a.getGenericPyramidalDefaults() == {"warping": False, "intermediateScaling": True, "scaling": False}
    assert (a.Niter, a.windowWidth, a.windowHeight, a.windowHalfWidth) == (5, 27, 27, 13)
    assert a._ofri_native_kind == "LK" and a.evaluateVorticityEnhancement(None, None) == [0, 0, 0, 0]
    p = a.native_params((0, 1, 0, 1))
    assert (p.n_iters, p.half_window, list(p.asym)) == (5, 13, [0, 1, 0, 1])
    e = LK.denseLucasKanade_PyCl(enableVorticityEnhancement=True)
    assert e._ofri_native_kind is None          # its window switches depend on every call's flow: host callback path
    yy, xx = np.mgrid[0:40, 0:50].astype(np.float32)
    rng = np.random.default_rng(0)
    for sign in (1.0, -1.0, 0.0):
        # solid-body rotation (u, v) = w (-(y - yc), x - xc): vorticity of a fixed sign, plus noise
        U = (-sign * 0.01 * (yy - 20) + rng.normal(0, 1e-3, yy.shape)).astype(np.float32)
        V = (sign * 0.01 * (xx - 25) + rng.normal(0, 1e-3, yy.shape)).astype(np.float32)
        got = e.evaluateVorticityEnhancement(U, V)
        assert got == LKO.vorticity_switch(U, V, True)
        assert got == {1.0: [0, 1, 0, 1], -1.0: [1, 0, 0, 1], 0.0: [0, 0, 0, 0]}[sign]
