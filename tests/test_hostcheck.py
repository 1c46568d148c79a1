"""CPU check of the CUDA kernels' per-pixel arithmetic.  tests/hostcheck/hostcheck.cpp includes the SAME header the
kernels include (opticalflow_ri_b200/csrc/ofri_pixel.cuh + ofri_tables.h), compiled with g++, and is compared with the
reference-generated golden vectors.  One-off stages must be bit-exact; the two iterative solvers use the fast
f32/FMA formulation and must stay within 2e-6 px of the reference on these cases.  Test-only: the product never
loads this harness."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ofri_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
OUT = os.path.join(HERE, "hostcheck", "_build", "libhostcheck.so")
_fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def hc():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", OUT], check=True)
    L = C.CDLL(OUT)
    L.hc_ls_iterate.restype = C.c_int
    return L


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def p(a):
    return a.ctypes.data_as(_fp)


GK = {"g34_3": (3.4, 3), "g048_5": (0.48, 5), "g12_7": (1.2, 7), "g18_9": (1.8, 9)}


@pytest.mark.parametrize("tag", sorted(GK))
def test_gauss(hc, stages, tag):
    i = f32(stages["gauss_in"])
    o = np.empty_like(i)
    k = f32(stages["gk_" + tag])
    hc.hc_gauss(p(i), i.shape[0], i.shape[1], p(k), len(k), p(o))
    assert np.array_equal(o, stages["gauss_out_" + tag])


@pytest.mark.parametrize("tag", list("abcdq"))
def test_resize(hc, stages, tag):
    i = f32(stages["rs_in_" + tag])
    ref = stages["rs_out_" + tag]
    o = np.empty_like(ref)
    hc.hc_resize(p(i), i.shape[0], i.shape[1], ref.shape[0], ref.shape[1], p(o))
    assert np.array_equal(o, ref)
    assert hc.hc_level_size(61, C.c_double(0.5)) == 30 and hc.hc_level_size(47, C.c_double(0.5)) == 24


@pytest.mark.parametrize("tag", list("abcd"))
@pytest.mark.parametrize("sc", [0, 1])
def test_spline_and_warp(hc, stages, tag, sc):
    s = "_s%d_" % sc
    Ua, Va = f32(stages["up_Ua_" + tag]), f32(stages["up_Va_" + tag])
    n1, n2 = f32(stages["up_n1_" + tag]), f32(stages["up_n2_" + tag])
    H, W = n1.shape
    h, w = Ua.shape
    mx = np.float32(np.float32(W) / np.float32(w)) if sc else np.float32(1)
    my = np.float32(np.float32(H) / np.float32(h)) if sc else np.float32(1)
    us, vs = np.empty((H, W), np.float32), np.empty((H, W), np.float32)
    hc.hc_spline(p(Ua), h, w, H, W, C.c_float(mx), p(us))
    hc.hc_spline(p(Va), h, w, H, W, C.c_float(my), p(vs))
    assert np.array_equal(us, stages["up_Uacc" + s + tag]) and np.array_equal(vs, stages["up_Vacc" + s + tag])
    o1, o2 = np.empty_like(n1), np.empty_like(n1)
    hc.hc_warp_pair(p(n1), p(n2), p(us), p(vs), H, W, p(o1), p(o2))
    assert np.array_equal(o1, stages["up_w1" + s + tag]) and np.array_equal(o2, stages["up_w2" + s + tag])


def test_spline_windowed_line_is_bit_identical(hc):
    """The chunk-parallel / windowed Thomas solve (ofri_spline.cuh: 48-row warm-up of both recurrences) reproduces the
    sequential full-line solve bit for bit in float64 -- every chunk size, whole lines and interior windows."""
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(0)
    for n in (4, 5, 7, 17, 64, 257, 1000):
        y = np.cumsum(rng.normal(size=n)) * 10.0 ** float(rng.integers(-3, 3))
        M = np.zeros(n)
        hc.hc_spline_line(y.ctypes.data_as(dp), n, M.ctypes.data_as(dp))
        for chunk in (2, 7, 33, 200):
            for a, b in ((0, n - 1), (0, min(n - 1, 3)), (max(0, n - 3), n - 1), (n // 3, min(n - 1, n // 3 + n // 4))):
                Mw = np.full(n, np.nan)
                hc.hc_spline_line_win(y.ctypes.data_as(dp), n, a, b, chunk, 48, Mw.ctypes.data_as(dp))
                assert np.array_equal(Mw[a:b + 1].view(np.uint64), M[a:b + 1].view(np.uint64)), (n, chunk, a, b)


@pytest.mark.parametrize("shape", [(4, 4, 8, 8), (5, 7, 11, 13), (150, 200, 300, 400), (257, 130, 513, 259)])
def test_spline_windowed_bands(hc, shape):
    """Rows [row0, row0 + rows) of the up-sampled plane from the window of coarse rows a band holds == the same rows of
    the whole-plane result (what the row-band mode relies on instead of an all-gather)."""
    h, w, H, W = shape
    rng = np.random.default_rng(1)
    a = (rng.normal(size=(h, w)) * 3).astype(np.float32)
    ref = np.empty((H, W), np.float32)
    hc.hc_spline(p(a), h, w, H, W, C.c_float(2.0), p(ref))
    for row0, rows in ((0, H), (0, min(H, 7)), (H // 3, H // 4 + 1), (max(H - 5, 0), min(5, H))):
        for cy, cx in ((2, 9), (57, 33)):
            out = np.empty((rows, W), np.float32)
            hc.hc_spline_win(p(a), h, w, H, W, C.c_float(2.0), row0, rows, cy, cx, 48, p(out))
            assert np.array_equal(out, ref[row0:row0 + rows]), (row0, rows, cy, cx)


def test_warp_coords(hc, stages):
    i = f32(stages["warp_img"])
    o = np.empty_like(i)
    hc.hc_warp_coords(p(i), p(f32(stages["warp_cy"])), p(f32(stages["warp_cx"])), i.shape[0], i.shape[1], p(o))
    assert np.array_equal(o, stages["warp_out"])


def test_hs(hc, stages):
    f1, f2 = f32(stages["hs_f1"]), f32(stages["hs_f2"])
    H, W = f1.shape
    fx, fy, ft = np.empty_like(f1), np.empty_like(f1), np.empty_like(f1)
    hc.hc_hs_derivs(p(f1), p(f2), H, W, p(fx), p(fy), p(ft))
    assert np.array_equal(fx, stages["hs_fx"]) and np.array_equal(fy, stages["hs_fy"])
    assert np.array_equal(ft, stages["hs_ft"])
    for nit in (1, 2, 7, 50):
        uo, vo = np.empty_like(f1), np.empty_like(f1)
        hc.hc_hs_iterate(p(f32(stages["hs_U0"])), p(f32(stages["hs_V0"])), p(fx), p(fy), p(ft), H, W, C.c_float(3.0),
                         nit, p(uo), p(vo))
        assert np.max(np.abs(uo - stages["hs_U_%d" % nit])) < 2e-6
        assert np.max(np.abs(vo - stages["hs_V_%d" % nit])) < 2e-6
        # "precise" formulation (reference arithmetic): bit-exact
        hc.hc_hs_iterate_precise(p(f32(stages["hs_U0"])), p(f32(stages["hs_V0"])), p(fx), p(fy), p(ft), H, W,
                                 C.c_float(3.0), nit, p(uo), p(vo))
        assert np.array_equal(uo, stages["hs_U_%d" % nit]) and np.array_equal(vo, stages["hs_V_%d" % nit])


def test_hs_precise_bom_row_bit_exact(hc, configs_small):
    """alpha = 1, 100 sweeps, no pre-filter (benchmark_of_methods.py row HS_Fs0_0): flows of +-26 px, bit for bit."""
    s = configs_small
    c0, c1 = f32(s["crop0"]), f32(s["crop1"])
    H, W = c0.shape
    fx, fy, ft = np.empty_like(c0), np.empty_like(c0), np.empty_like(c0)
    hc.hc_hs_derivs(p(c0), p(c1), H, W, p(fx), p(fy), p(ft))
    z = np.zeros_like(c0)
    uo, vo = np.empty_like(c0), np.empty_like(c0)
    hc.hc_hs_iterate_precise(p(z), p(z), p(fx), p(fy), p(ft), H, W, C.c_float(1.0), 100, p(uo), p(vo))
    assert np.array_equal(uo, s["bom_HS_Fs0_0_U"]) and np.array_equal(vo, s["bom_HS_Fs0_0_V"])


@pytest.mark.parametrize("tag,h", [("h5", 5), ("h01", 0.1)])
def test_ls(hc, stages, tag, h):
    g1, g2 = f32(stages["ls_g1"]), f32(stages["ls_g2"])
    H, W = g1.shape
    coef = np.empty((8, H, W), np.float32)
    hc.hc_ls_coef(p(g1), p(g2), H, W, C.c_float(h), p(coef))
    ref = O.ls_coefficients(g1, g2, h)          # the oracle is itself pinned bit-exactly to the reference
    for i in range(8):
        assert np.array_equal(coef[i], ref[i]), i
    assert np.array_equal(coef[5], stages["ls_B11_" + tag])
    uo, vo = np.empty_like(g1), np.empty_like(g1)
    err = C.c_double()
    k = hc.hc_ls_iterate(p(f32(stages["ls_Vin"])), p(f32(stages["ls_Uin"])), p(coef), H, W, C.c_float(h), 60,
                         C.c_double(1e-8), p(uo), p(vo), C.byref(err))
    assert k == 60
    assert np.max(np.abs(vo - stages["ls_U_" + tag])) < 1e-6 and np.max(np.abs(uo - stages["ls_V_" + tag])) < 1e-6
    assert err.value == pytest.approx(float(stages["ls_err_" + tag]), rel=1e-4)
