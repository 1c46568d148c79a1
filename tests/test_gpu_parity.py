"""GPU parity tests (run on the B200 box with `-m gpu`).  Every call goes through the C ABI of libofri.so
(opticalflow_ri_b200.Handle is a thin ctypes binding) and is compared with

  * the golden vectors produced by the unmodified reference (tests/golden/*.npz), and
  * the CPU oracle (oracle/ofri_oracle.py, itself bit-exact against those vectors) on seeded inputs.

Tolerances (north_star): one-off stages bit-exact; flows max|dU|, |dV| <= 1e-4 px; EPE-RMSE difference <= 1e-6 px.
The fused (temporally blocked) kernels must be BIT-IDENTICAL to the one-sweep-per-launch kernels."""
import os
import sys

import numpy as np
import pytest

import ofri_oracle as O

pytestmark = pytest.mark.gpu

TOL_FLOW = 1e-4
TOL_EPE = 1e-6


@pytest.fixture(scope="module")
def ofri():
    import opticalflow_ri_b200 as o
    return o


@pytest.fixture(scope="module")
def h(ofri):
    hd = ofri.Handle(0)          # raises (no CPU fallback) if the CUDA library / GPU is missing
    hd.set_option("auto_fuse", 0)   # the tests below pick fuse factors explicitly (see test_auto_fuse_is_invisible)
    yield hd
    hd.close()


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        raise AssertionError("not bit-exact: max|d| = %g at %s (%d of %d differ)" %
                             (d.max(), np.unravel_index(np.argmax(d), d.shape), np.count_nonzero(a != b), a.size))


def close(a, b, tol):
    d = np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))
    assert d <= tol, "max|d| = %g > %g" % (d, tol)


def rand_img(rng, H, W):
    img = rng.uniform(0, 12, (H, W))
    n = max(4, H * W // 40)
    img[rng.integers(0, H, n), rng.integers(0, W, n)] += rng.uniform(40, 240, n)
    return np.clip(np.rint(img), 0, 255).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------
# stages vs reference goldens (bit-exact)
# ---------------------------------------------------------------------------------------------------------------
GK = {"g34_3": (3.4, 3), "g048_5": (0.48, 5), "g12_7": (1.2, 7), "g18_9": (1.8, 9)}


@pytest.mark.parametrize("tag", sorted(GK))
def test_gauss_golden(h, ofri, stages, tag):
    same(h.gauss_px(stages["gauss_in"], ofri.gaussian_taps(*GK[tag])), stages["gauss_out_" + tag])


def test_gauss_tiny_and_batched(h, ofri, stages):
    same(h.gauss_px(stages["gauss_small_in"], ofri.gaussian_taps(0.48, 5)), stages["gauss_small_out"])
    rng = np.random.default_rng(3)
    imgs = np.stack([rand_img(rng, 67, 131) for _ in range(5)])
    for sg, K in ((3.4, 3), (0.48, 5), (1.8, 9)):
        out = h.gauss_px(imgs, ofri.gaussian_taps(sg, K))
        for b in range(5):
            same(out[b], O.gaussian_filter_px(imgs[b], sg, K))


@pytest.mark.parametrize("tag", list("abcdq"))
def test_resize_golden(h, stages, tag):
    ref = stages["rs_out_" + tag]
    same(h.resize_bicubic(stages["rs_in_" + tag], ref.shape[0], ref.shape[1]), ref)


def test_resize_oracle_sizes(h):
    rng = np.random.default_rng(4)
    for H, W in ((512, 512), (301, 517), (96, 1024)):
        img = rand_img(rng, H, W)
        oh, ow = h.level_size(H, 0.5), h.level_size(W, 0.5)
        assert (oh, ow) == (O.level_size(H, 0.5), O.level_size(W, 0.5))
        same(h.resize_bicubic(img, oh, ow), O.imresize_bicubic(img, ow, oh))


@pytest.mark.parametrize("tag", list("abcd"))
@pytest.mark.parametrize("sc", [0, 1])
def test_spline_warp_golden(h, stages, tag, sc):
    s = "_s%d_" % sc
    Ua, Va = stages["up_Ua_" + tag], stages["up_Va_" + tag]
    n1, n2 = stages["up_n1_" + tag], stages["up_n2_" + tag]
    H, W = n1.shape
    hh, ww = Ua.shape
    mx = np.float32(np.float32(W) / np.float32(ww)) if sc else 1.0
    my = np.float32(np.float32(H) / np.float32(hh)) if sc else 1.0
    us = h.spline_upsample(Ua, H, W, mx)
    vs = h.spline_upsample(Va, H, W, my)
    same(us, stages["up_Uacc" + s + tag])
    same(vs, stages["up_Vacc" + s + tag])
    w1, w2 = h.warp_pair(n1, n2, us, vs)
    same(w1, stages["up_w1" + s + tag])
    same(w2, stages["up_w2" + s + tag])


def test_spline_oracle_big(h):
    rng = np.random.default_rng(5)
    a = rng.normal(0, 2, (3, 150, 259)).astype(np.float32)
    out = h.spline_upsample(a, 301, 517, 2.0)
    for b in range(3):
        same(out[b], (O.spline_upsample(a[b], 301, 517) * np.float32(2.0)).astype(np.float32))


@pytest.mark.parametrize("shape", [(4, 4, 8, 9), (37, 51, 75, 101), (512, 512, 1024, 1024), (1000, 700, 2000, 1400),
                                   (9, 14000, 18, 28000), (12, 9000, 25, 19999), (2100, 40, 4200, 80)])
def test_spline_chunked_equals_sequential(h, shape):
    """spline_variant 1 (chunk-parallel windowed column solve + fused per-row kernel, the default) against variant 0
    (one thread per line, sequential full-length Thomas solves): bit-identical float32 planes, including widths that
    need several x segments in shared memory (> 4252 knots per segment) and columns long enough for many chunks."""
    hh, ww, H, W = shape
    rng = np.random.default_rng(hh + ww)
    a = (np.cumsum(rng.normal(0, 0.3, (2, hh, ww)), axis=2) + rng.normal(0, 1, (2, hh, ww))).astype(np.float32)
    try:
        h.set_option("spline_variant", 0)
        ref = h.spline_upsample(a, H, W, 2.0)
    finally:
        h.set_option("spline_variant", 1)
    out = h.spline_upsample(a, H, W, 2.0)
    same(out, ref)


def test_warp_coords_golden(h, stages):
    same(h.warp_bilinear(stages["warp_img"], stages["warp_cy"], stages["warp_cx"]), stages["warp_out"])


def test_spline_too_small_raises(h):
    with pytest.raises(ValueError):
        h.spline_upsample(np.zeros((3, 8), np.float32), 6, 16)


def test_hs_derivatives_golden(h, stages):
    fx, fy, ft = h.hs_derivatives(stages["hs_f1"], stages["hs_f2"])
    same(fx, stages["hs_fx"])
    same(fy, stages["hs_fy"])
    same(ft, stages["hs_ft"])


@pytest.mark.parametrize("tag,hp", [("h5", 5), ("h01", 0.1)])
def test_ls_coefficients_golden(h, stages, tag, hp):
    coef = h.ls_coefficients(stages["ls_g1"], stages["ls_g2"], hp)
    ref = O.ls_coefficients(stages["ls_g1"], stages["ls_g2"], hp)
    for i in range(8):
        same(coef[i], ref[i])
    same(coef[5], stages["ls_B11_" + tag])
    same(coef[6], stages["ls_B12_" + tag])
    same(coef[7], stages["ls_B22_" + tag])
    if tag == "h5":
        same(coef[3], stages["ls_Ixt"])
        same(coef[4], stages["ls_Iyt"])


# ---------------------------------------------------------------------------------------------------------------
# Horn-Schunck iterations
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nit", [1, 2, 7, 50])
@pytest.mark.parametrize("fuse", [0, 4])
def test_hs_compute_golden(h, stages, nit, fuse):
    h.set_option("hs_fuse", fuse)
    try:
        h.set_option("hs_precise", 0)      # fast f32/FMA formulation: within 2e-6 px
        U, V, err = h.hs_compute(stages["hs_f1"], stages["hs_f2"], stages["hs_U0"], stages["hs_V0"], 3.0, nit)
        close(U, stages["hs_U_%d" % nit], 2e-6)
        close(V, stages["hs_V_%d" % nit], 2e-6)
        assert err == pytest.approx(float(stages["hs_err_%d" % nit]), rel=1e-4)
        h.set_option("hs_precise", 2)      # reference arithmetic: bit-exact
        U, V, err = h.hs_compute(stages["hs_f1"], stages["hs_f2"], stages["hs_U0"], stages["hs_V0"], 3.0, nit)
        same(U, stages["hs_U_%d" % nit])
        same(V, stages["hs_V_%d" % nit])
    finally:
        h.set_option("hs_fuse", 4)
        h.set_option("hs_precise", 1)


@pytest.mark.parametrize("shape", [(45, 58), (301, 517), (512, 512), (2, 2), (9, 1030), (122, 130), (247, 120)])
def test_hs_fused_bit_identical_to_simple(h, shape):
    """Temporal blocking must not change a single bit: every (T, tile variant) against one-sweep-per-launch."""
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    H, W = shape
    B = 3
    fx = rng.normal(0, 8, (B, H, W)).astype(np.float32)
    fy = rng.normal(0, 8, (B, H, W)).astype(np.float32)
    ft = rng.normal(0, 8, (B, H, W)).astype(np.float32)
    U0 = rng.normal(0, 1, (B, H, W)).astype(np.float32)
    V0 = rng.normal(0, 1, (B, H, W)).astype(np.float32)
    nit = 13
    try:
        for precise in (0, 2):
            h.set_option("hs_precise", precise)
            h.set_option("hs_fuse", 0)
            Ur, Vr = h.hs_iterate(U0, V0, fx, fy, ft, 7.5, nit)
            for T in (1, 2, 3, 4, 5, 6, 8):
                for variant in (0, 24):      # shared-memory kernel / persistent TMA kernel
                    h.set_option("hs_fuse", T)
                    h.set_option("hs_variant", variant)
                    U, V = h.hs_iterate(U0, V0, fx, fy, ft, 7.5, nit)
                    try:
                        same(U, Ur)
                        same(V, Vr)
                    except AssertionError as e:
                        raise AssertionError("precise=%d T=%d variant=%d shape=%s: %s" % (precise, T, variant, shape, e))
    finally:
        h.set_option("hs_fuse", 4)
        h.set_option("hs_variant", 24)
        h.set_option("hs_precise", 1)


def test_hs_simple_vs_oracle(h):
    rng = np.random.default_rng(11)
    f1 = O.gaussian_filter_px(rand_img(rng, 120, 97), 3.4, 3)
    f2 = O.gaussian_filter_px(rand_img(rng, 120, 97), 3.4, 3)
    Uo, Vo, eo = O.hs_compute(f1, f2, 21.0, 40, np.zeros_like(f1), np.zeros_like(f1))
    try:
        h.set_option("hs_precise", 0)
        U, V, err = h.hs_compute(f1, f2, None, None, 21.0, 40)
        close(U, Uo, 2e-6)
        close(V, Vo, 2e-6)
        assert err == pytest.approx(eo, rel=1e-4)
        h.set_option("hs_precise", 2)
        U, V, err = h.hs_compute(f1, f2, None, None, 21.0, 40)
        same(U, Uo)
        same(V, Vo)
    finally:
        h.set_option("hs_precise", 1)


# ---------------------------------------------------------------------------------------------------------------
# Liu-Shen
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,hp", [("h5", 5), ("h01", 0.1)])
@pytest.mark.parametrize("fuse", [0, 1, 2, 3, 4])
def test_ls_compute_golden(h, stages, tag, hp, fuse):
    h.set_option("ls_fuse", fuse)
    try:
        U, V, err, it = h.ls_compute(stages["ls_g1"], stages["ls_g2"], stages["ls_Uin"], stages["ls_Vin"], hp)
    finally:
        h.set_option("ls_fuse", 2)
    assert it == 60
    close(U, stages["ls_U_" + tag], 1e-6)
    close(V, stages["ls_V_" + tag], 1e-6)
    assert err == pytest.approx(float(stages["ls_err_" + tag]), rel=1e-4)


def test_ls_first_sweep_golden(h, stages):
    U, V, err, it = h.ls_compute(stages["ls_g1"], stages["ls_g2"], stages["ls_Uin"], stages["ls_Vin"], 5, maxiter=1)
    assert it == 1
    close(V, stages["ls_u1_h5"], 2e-7)     # inside the solver u is the ROW component = our V
    close(U, stages["ls_v1_h5"], 2e-7)
    assert err == pytest.approx(float(stages["ls_err0_h5"]), rel=1e-5)


@pytest.mark.parametrize("fuse", [0, 1, 2, 3, 4])
def test_ls_early_exit_identical_frames(h, stages, fuse):
    """total_error == 0 after the first sweep -> the reference stops after ONE sweep (LS:141)."""
    h.set_option("ls_fuse", fuse)
    try:
        U, V, err, it = h.ls_compute(stages["ls_g1"], stages["ls_g1"], None, None, 5)
    finally:
        h.set_option("ls_fuse", 2)
    assert it == int(stages["ls_same_niter"]) == 1 and err == 0.0
    same(U, stages["ls_same_U"])
    same(V, stages["ls_same_V"])


@pytest.mark.parametrize("shape", [(40, 52), (151, 259), (256, 256), (2, 3)])
def test_ls_fused_bit_identical_and_midblock_stop(h, shape):
    """Fused blocks vs single sweeps, including tolerances that trip in the MIDDLE of a fused block (replay path)."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    H, W = shape
    B = 3
    g1 = np.stack([O.gaussian_filter_px(rand_img(rng, H, W), 0.48, 5) if min(H, W) >= 2 else rand_img(rng, H, W)
                   for _ in range(B)])
    g2 = np.stack([O.gaussian_filter_px(rand_img(rng, H, W), 0.48, 5) if min(H, W) >= 2 else rand_img(rng, H, W)
                   for _ in range(B)])
    U0 = rng.normal(0, 0.3, (B, H, W)).astype(np.float32)
    V0 = rng.normal(0, 0.3, (B, H, W)).astype(np.float32)
    try:
        for maxiter, tol in ((60, 1e-8), (17, 1e-8), (60, 2e-4), (60, 5e-5), (60, 1e-3)):
            h.set_option("ls_fuse", 0)
            Ur, Vr, er, itr = h.ls_compute(g1, g2, U0, V0, 5, maxiter=maxiter, tol=tol)
            for T in (1, 2, 3, 4):
                for lv in (0, 8):            # shared-memory kernel / persistent TMA kernel
                    h.set_option("ls_fuse", T)
                    h.set_option("ls_variant", lv)
                    U, V, e, it = h.ls_compute(g1, g2, U0, V0, 5, maxiter=maxiter, tol=tol)
                    assert list(it) == list(itr), (T, lv, maxiter, tol, list(it), list(itr))
                    try:
                        same(U, Ur)
                        same(V, Vr)
                    except AssertionError as ex:
                        raise AssertionError("T=%d variant=%d maxiter=%d tol=%g its=%s: %s" % (T, lv, maxiter, tol,
                                                                                                   list(itr), ex))
                    np.testing.assert_allclose(e, er, rtol=1e-5)
    finally:
        h.set_option("ls_fuse", 2)
        h.set_option("ls_variant", 8)


def test_ls_stop_rule_vs_oracle(h):
    rng = np.random.default_rng(21)
    g1 = O.gaussian_filter_px(rand_img(rng, 64, 80), 0.48, 5)
    g2 = O.gaussian_filter_px(rand_img(rng, 64, 80), 0.48, 5)
    for tol in (1e-8, 1e-4, 3e-5):
        Uo, Vo, eo, ko = O.ls_compute(g1, g2, 5, np.zeros_like(g1), np.zeros_like(g1), 60, tol)
        U, V, err, it = h.ls_compute(g1, g2, None, None, 5, maxiter=60, tol=tol)
        assert it == ko, (tol, it, ko)
        close(U, Uo, 1e-6)
        close(V, Vo, 1e-6)
        assert err == pytest.approx(eo, rel=1e-4)


# ---------------------------------------------------------------------------------------------------------------
# whole driver vs reference goldens
# ---------------------------------------------------------------------------------------------------------------
def flow(h, ofri, im1, im2, FILTER, main, L=1, K=1, FILTER_OPT=None, opt=None, **kw):
    p = ofri.make_params(main, opt, filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT, pyramid_levels=L, k_levels=K, **kw)
    return h.pyramidal_flow(im1, im2, p)


HS_DEF = dict(warping=True, bilinear=True, final_scaling=True)     # HS adapter defaults (HornSchunck.py:45-50)


def test_driver_bom_rows(h, ofri, configs_small):
    s = configs_small
    for name, fs, L, use_ls in (("HS_Fs0_0", 0.0, 1, False), ("HS_Fs3_4", 3.4, 1, False),
                                ("HS_Fs3_4_PyrLvls2", 3.4, 2, False), ("LiuSE_HS_Fs3_4_PyrLvls2", 3.4, 2, True)):
        if use_ls:
            U, V = flow(h, ofri, s["crop0"], s["crop1"], fs, ofri.ls_algo(0.1), L)
        else:
            U, V = flow(h, ofri, s["crop0"], s["crop1"], fs, ofri.hs_algo([1.0] * L, 100), L, **HS_DEF)
        close(U, s["bom_%s_U" % name], TOL_FLOW)
        close(V, s["bom_%s_V" % name], TOL_FLOW)


def test_driver_variants(h, ofri, configs_small):
    s = configs_small
    c0, c1 = s["crop0"], s["crop1"]
    cases = {
        "odd_c3": lambda: flow(h, ofri, s["odd0"], s["odd1"], 3.4, ofri.hs_algo([45, 21], 200), 2, 1, 0.48,
                               ofri.ls_algo(5), **HS_DEF),
        "l3": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 30, 21], 150), 3, 1, 0.48, ofri.ls_algo(5), **HS_DEF),
        "k2": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 45, 21, 21], 100), 2, 2, 0.48, ofri.ls_algo(5),
                           **HS_DEF),
        "k2f08": lambda: flow(h, ofri, c0, c1, 0.8, ofri.hs_algo([45, 45, 21, 21], 100), 2, 2, **HS_DEF),
        "nowarp": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 21], 100), 2, 1, warping=False,
                               intermediate_scaling=True, final_scaling=True),
        "nowarp_k2": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 45, 21, 21], 60), 2, 2, warping=False,
                                  final_scaling=True),
        "lsmain_hsopt": lambda: flow(h, ofri, c0, c1, 3.4, ofri.ls_algo(5), 2, 1, 0.0, ofri.hs_algo([30, 30], 40)),
    }
    bad = []
    for name, fn in cases.items():
        U, V = fn()
        du = np.max(np.abs(U - s[name + "_U"]))
        dv = np.max(np.abs(V - s[name + "_V"]))
        if not (du <= TOL_FLOW and dv <= TOL_FLOW):
            bad.append((name, float(du), float(dv)))
    assert not bad, bad


@pytest.mark.parametrize("cfg", ["c1", "c2", "c3"])
def test_driver_baseline_configs_bundled(h, ofri, configs_bundled, bundled_pair, cfg):
    """BASELINE.json configs 1-3 on the bundled Poiseuille pair: flow within 1e-4 px of the reference and EPE-RMSE
    (against the analytic Poiseuille profile) within 1e-6 px of the reference's."""
    I0, I1 = bundled_pair
    if cfg == "c1":
        U, V = flow(h, ofri, I0, I1, 3.4, ofri.hs_algo([21], 600), 1, **HS_DEF)
    elif cfg == "c2":
        U, V = flow(h, ofri, I0, I1, 3.4, ofri.hs_algo([45, 21], 600), 2, **HS_DEF)
    else:
        U, V = flow(h, ofri, I0, I1, 3.4, ofri.hs_algo([45, 21], 600), 2, 1, 0.48, ofri.ls_algo(5), **HS_DEF)
    Ur, Vr = configs_bundled[cfg + "_U"], configs_bundled[cfg + "_V"]
    du, dv = np.max(np.abs(U - Ur)), np.max(np.abs(V - Vr))
    Ut, Vt = O.poiseuille_truth(512, 512)
    de = abs(O.epe_rmse(U, V, Ut, Vt) - O.epe_rmse(Ur, Vr, Ut, Vt))
    print("\n%s: max|dU| %.3g max|dV| %.3g |dEPE-RMSE| %.3g" % (cfg, du, dv, de))
    assert du <= TOL_FLOW and dv <= TOL_FLOW, (du, dv)
    assert de <= TOL_EPE, de


@pytest.mark.parametrize("alpha", [1.0, 5.0, 21.0])
def test_hs_precise_rule_liu_shen_contracts_coarse_rounding(h, ofri, bundled_pair, alpha):
    """hs_precise = 1 (default) runs a coarse level's Horn-Schunck solve in the fast arithmetic when the Liu-Shen
    refinement follows on the same level, and in the reference arithmetic when its result reaches the warp as it is.
    Against the reference arithmetic on every level (hs_precise = 2): <= 2e-5 px either way on the bundled pair (a
    fifth of the 1e-4 px bar), whereas the fast arithmetic everywhere WITHOUT the refinement is off by more than the bar
    for weak regularisation -- which is why the rule exists."""
    I0, I1 = bundled_pair
    mk_ls = lambda: ofri.make_params(ofri.hs_algo([alpha, alpha], 200), ofri.ls_algo(5), filter_sigma=3.4,
                                     filter_opt_sigma=0.48, pyramid_levels=2, **HS_DEF)
    mk_hs = lambda: ofri.make_params(ofri.hs_algo([alpha, alpha], 200), filter_sigma=3.4, pyramid_levels=2, **HS_DEF)
    try:
        res = {}
        for mode in (2, 1, 0):
            h.set_option("hs_precise", mode)
            h.set_option("timing", 1)
            res["ls", mode] = h.pyramidal_flow(I0, I1, mk_ls())
            stages_ls = dict(h.stage_timings())
            res["hs", mode] = h.pyramidal_flow(I0, I1, mk_hs())
            stages_hs = dict(h.stage_timings())
            if mode == 1:      # which kernels ran: no precise sweeps with the refinement, precise coarse level without
                assert stages_ls.get("hs_iterate_precise", 0.0) == 0.0 and stages_ls.get("hs_iterate", 0.0) > 0.0
                assert stages_hs.get("hs_iterate_precise", 0.0) > 0.0
        dev = lambda k, m: max(float(np.max(np.abs(res[k, m][i] - res[k, 2][i]))) for i in (0, 1))
        print("\nalpha %g: HS+LS default %.3g, HS default %.3g, HS all-fast %.3g px" % (alpha, dev("ls", 1), dev("hs", 1),
                                                                                      dev("hs", 0)))
        assert dev("ls", 1) <= 2e-5 and dev("hs", 1) <= 2e-5
        if alpha <= 1.0:
            assert dev("hs", 0) > TOL_FLOW
    finally:
        h.set_option("hs_precise", 1)
        h.set_option("timing", 0)


def test_auto_fuse_is_invisible(h, ofri, configs_small):
    """Launches that cannot fill the GPU fuse deeper (8 HS / 4 LS sweeps per launch): same bits."""
    s = configs_small
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 52), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    U0, V0 = h.pyramidal_flow(s["crop0"], s["crop1"], mk())
    try:
        h.set_option("auto_fuse", 1)
        U1, V1 = h.pyramidal_flow(s["crop0"], s["crop1"], mk())
    finally:
        h.set_option("auto_fuse", 0)
    same(U1, U0)
    same(V1, V0)


def test_errors_reported(h, ofri, configs_small):
    s = configs_small
    p = ofri.make_params(ofri.hs_algo([45, 21], 50), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                         pyramid_levels=2, **HS_DEF)
    U, V, err = h.pyramidal_flow(s["crop0"], s["crop1"], p, want_errors=True)
    assert err.shape == (2, 2) and np.all(np.isfinite(err)) and np.all(err > 0)
    tr = []
    O.pyramidal_flow(s["crop0"], s["crop1"], 3.4, O.HSParams([21, 45], 50), 2, 1, 0.48, O.LSParams(5), trace=tr)
    for i in range(2):
        assert err[i, 0] == pytest.approx(tr[i]["err_main"], rel=1e-3)
        assert err[i, 1] == pytest.approx(tr[i]["err_opt"], rel=1e-3)


def test_batch_equals_single_and_chunking(h, ofri, configs_small):
    s = configs_small
    a = np.stack([s["crop0"], s["crop1"], s["crop0"][::-1].copy(), s["crop1"][:, ::-1].copy(), s["crop0"]])
    b = np.stack([s["crop1"], s["crop0"], s["crop1"][::-1].copy(), s["crop0"][:, ::-1].copy(), s["crop0"]])
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 60), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    U, V = h.pyramidal_flow(a, b, mk())
    for i in range(5):
        Ui, Vi = h.pyramidal_flow(a[i], b[i], mk())
        same(U[i], Ui)
        same(V[i], Vi)
    try:
        h.set_option("chunk_pairs", 2)          # 3 chunks through the double-buffered staging path
        U2, V2 = h.pyramidal_flow(a, b, mk())
    finally:
        h.set_option("chunk_pairs", 0)
    same(U2, U)
    same(V2, V)
    assert np.all(U[4] == 0) and np.all(V[4] == 0)      # identical frames -> zero flow


def test_error_codes(h, ofri):
    z = np.zeros((16, 16), np.float32)
    with pytest.raises(IndexError):
        h.pyramidal_flow(z, z, ofri.make_params(ofri.hs_algo([21], 5), pyramid_levels=2, **HS_DEF))
    with pytest.raises(ValueError):       # 16 -> 2 px at level 1 of 4: too small for the cubic spline
        h.pyramidal_flow(z, z, ofri.make_params(ofri.hs_algo([1, 1, 1, 1], 5), pyramid_levels=4, **HS_DEF))
    with pytest.raises(ValueError):       # Liu-Shen warp: the 73-tap Gaussian does not fit a 16 x 16 level
        h.pyramidal_flow(z, z, ofri.make_params(ofri.ls_algo(5), pyramid_levels=2, bilinear=False))


def test_synthetic_piv_vs_oracle_and_truth(h, ofri):
    """Seeded synthetic PIV pair with known Poiseuille truth (the bench workload's generator), odd size."""
    I0, I1 = O.synthetic_piv_pair(300, 420, seed=7)
    Uo, Vo = O.pyramidal_flow(I0, I1, 3.4, O.HSParams([21, 45], 150), 2, 1, 0.48, O.LSParams(5))
    U, V = flow(h, ofri, I0, I1, 3.4, ofri.hs_algo([45, 21], 150), 2, 1, 0.48, ofri.ls_algo(5), **HS_DEF)
    close(U, Uo, TOL_FLOW)
    close(V, Vo, TOL_FLOW)
    Ut, Vt = O.poiseuille_truth(300, 420)
    assert abs(O.epe_rmse(U, V, Ut, Vt) - O.epe_rmse(Uo, Vo, Ut, Vt)) <= TOL_EPE
    assert O.epe_rmse(U, V, Ut, Vt) < 0.6


def test_full_size_properties_1024(h, ofri):
    """BASELINE config 4's frame size (1024 x 1024), too slow for the oracle at full iteration count: use
    size-independent properties instead -- batch/chunk invariance, the x-mirror symmetry of the whole pipeline's
    index logic is NOT exact (asymmetric Gaussian padding), so check: (a) identical pairs give identical flows,
    (b) the flow of a pure Poiseuille pair recovers the profile, (c) fused == unfused bit-for-bit end to end."""
    pairs = [O.synthetic_piv_pair(1024, 1024, seed=s) for s in (0, 1)]
    a = np.stack([pairs[0][0], pairs[1][0], pairs[0][0]])
    b = np.stack([pairs[0][1], pairs[1][1], pairs[0][1]])
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 600), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    U, V = h.pyramidal_flow(a, b, mk())
    same(U[0], U[2])
    same(V[0], V[2])
    Ut, Vt = O.poiseuille_truth(1024, 1024)
    for i in range(2):
        assert O.epe_rmse(U[i], V[i], Ut, Vt) < 0.1
    try:
        h.set_option("hs_fuse", 0)
        h.set_option("ls_fuse", 0)
        U0, V0 = h.pyramidal_flow(a[:1], b[:1], mk())
    finally:
        h.set_option("hs_fuse", 4)
        h.set_option("ls_fuse", 2)
    same(U0[0], U[0])
    same(V0[0], V[0])
    try:          # all-levels reference arithmetic vs the default (coarse levels only): final level differs by rounding only
        h.set_option("hs_precise", 2)
        U2, V2 = h.pyramidal_flow(a[:1], b[:1], mk())
    finally:
        h.set_option("hs_precise", 1)
    close(U2[0], U[0], 2e-5)
    close(V2[0], V[0], 2e-5)


def test_reference_golden_1024_config4(h, ofri, big1024):
    """BASELINE config 4's frame size against the REFERENCE itself: one seeded synthetic 1024 x 1024 pair, full EX3
    parameters (HS 600 sweeps alphas [21, 45] + Liu-Shen h = 5, 2 levels), default options -- alone and as a member of a
    64-pair chunk (the launch shape of the benchmark: 9 x 18 tiles per pair, border tiles in x and y)."""
    g = big1024
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 600), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    Ut, Vt = O.poiseuille_truth(1024, 1024)
    e_ref = O.epe_rmse(g["U"], g["V"], Ut, Vt)
    hd = ofri.Handle(0)                     # a fresh handle: every option at its default (auto_fuse included)
    try:
        U, V = hd.pyramidal_flow(g["im0"], g["im1"], mk())
        du, dv = float(np.max(np.abs(U - g["U"]))), float(np.max(np.abs(V - g["V"])))
        de = abs(O.epe_rmse(U, V, Ut, Vt) - e_ref)
        print("\n1024^2 single vs reference: max|dU| %.3g max|dV| %.3g |dEPE-RMSE| %.3g (EPE-RMSE ref %.4f)" % (du, dv, de, e_ref))
        assert du <= TOL_FLOW and dv <= TOL_FLOW and de <= TOL_EPE, (du, dv, de)
        import torch                         # device memory only (the bench's device-pointer call, 64 pairs per chunk)
        other = O.synthetic_piv_pair(1024, 1024, seed=3)
        a = torch.from_numpy(np.stack([other[0]] * 64)).cuda()
        b = torch.from_numpy(np.stack([other[1]] * 64)).cuda()
        a[37], b[37] = torch.from_numpy(g["im0"]).cuda(), torch.from_numpy(g["im1"]).cuda()
        ub, vb = torch.empty_like(a), torch.empty_like(a)
        torch.cuda.synchronize()
        hd.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), 64, 1024, 1024, mk(), ub.data_ptr(), vb.data_ptr(), None, device=True)
        hd.synchronize()
        assert hd.get_option("last_chunk_pairs") == 64
        Ub, Vb = ub[37].cpu().numpy(), vb[37].cpu().numpy()
        du, dv = float(np.max(np.abs(Ub - g["U"]))), float(np.max(np.abs(Vb - g["V"])))
        print("1024^2 in a 64-pair chunk vs reference: max|dU| %.3g max|dV| %.3g" % (du, dv))
        assert du <= TOL_FLOW and dv <= TOL_FLOW, (du, dv)
        same(Ub, U)                         # and bit-identical to the single-pair call
        same(Vb, V)
    finally:
        hd.close()


def test_k_loop_weak_regularisation_goldens(h, ofri, configs_extra):
    """kLevels = 2 with alpha = 1 and no Liu-Shen refinement: every Horn-Schunck result but the last is consumed by a
    re-warp (GPOF:392-404), also on the LAST level, so those solves must run the reference arithmetic
    (hs_needs_precise / feeds_warp in run_pyramid)."""
    s = configs_extra
    c0, c1 = s["crop0"], s["crop1"]
    cases = {
        "k2a1": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([1.0] * 4, 100), 2, 2, **HS_DEF),
        "l1k2a1": lambda: flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([1.0] * 2, 100), 1, 2, **HS_DEF),
        "l3k2": lambda: flow(h, ofri, c0, c1, 2.0, ofri.hs_algo([2.0, 2.0, 2.0, 2.0, 0.5, 0.5], 60), 3, 2, **HS_DEF),
    }
    bad = []
    for name, fn in cases.items():
        U, V = fn()
        du = float(np.max(np.abs(U - s[name + "_U"])))
        dv = float(np.max(np.abs(V - s[name + "_V"])))
        print("\n%s: max|dU| %.3g max|dV| %.3g" % (name, du, dv))
        if not (du <= TOL_FLOW and dv <= TOL_FLOW):
            bad.append((name, du, dv))
    assert not bad, bad


# ---------------------------------------------------------------------------------------------------------------
# biLinear = False: the "Liu-Shen warp" branch (GPOF:190-196, 204-221)
# ---------------------------------------------------------------------------------------------------------------
def test_liu_shen_warp_stage_golden(h, lswarp):
    g = lswarp
    H, W = g["crop0"].shape
    us = h.spline_upsample(g["st_Ua"], H, W, np.float32(np.float32(W) / np.float32(g["st_Ua"].shape[1])))
    vs = h.spline_upsample(g["st_Va"], H, W, np.float32(np.float32(H) / np.float32(g["st_Va"].shape[0])))
    same(us, g["st_ua"])
    same(vs, g["st_va"])
    same(h.liu_shen_warp(g["crop0"], us, vs), g["st_w1"])       # scatter (wrap-around, collisions) + 73-tap Gaussian + OF equation
    bad = us.copy()
    bad[150, W - 1] = 0.7                                        # pushes a pixel past the right border
    with pytest.raises(IndexError):
        h.liu_shen_warp(g["crop0"], bad, vs)


def test_liu_shen_warp_driver_goldens(h, ofri, lswarp):
    g = lswarp
    c0, c1 = g["crop0"], g["crop1"]
    keep = c0.copy()
    U, V = flow(h, ofri, c0, c1, 3.4, ofri.ls_algo(0.1), 2, bilinear=False)
    close(U, g["lsmain_U"], TOL_FLOW)
    close(V, g["lsmain_V"], TOL_FLOW)
    U, V = flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 21], 100), 2, 1, 0.48, ofri.ls_algo(5), bilinear=False,
                final_scaling=True)
    close(U, g["hsnodef_U"], TOL_FLOW)
    close(V, g["hsnodef_V"], TOL_FLOW)
    U, V = flow(h, ofri, c0, c1, 3.4, ofri.ls_algo(1.0), 3, bilinear=False)
    close(U, g["l3_U"], TOL_FLOW)
    close(V, g["l3_V"], TOL_FLOW)
    assert int(g["k2_raises_index_error"]) == 1
    with pytest.raises(IndexError):           # the reference raises here too (a target leaves the frame at the bottom)
        flow(h, ofri, c0, c1, 3.4, ofri.hs_algo([45, 45, 21, 21], 60), 2, 2, bilinear=False, final_scaling=True)
    same(c0, keep)                            # unlike the reference, the caller's frame is left alone


# ---------------------------------------------------------------------------------------------------------------
# drop-in modules: the reference's example / benchmark call shapes, unchanged
# ---------------------------------------------------------------------------------------------------------------
def test_dropin_example3_and_wrapper(ofri, configs_bundled, bundled_pair, configs_small):
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow
        from GenericPyramidalOpticalFlowWrapper import GenericPyramidalOpticalFlowWrapper
        from HornSchunck import HSOpticalFlowAlgoAdapter
        from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter
        import gaussian_filter as GF
    finally:
        sys.path.remove(ofri.SRC_DIR)
    Iold, Inew = bundled_pair
    # examples/LiuSE_PyHSchunck_Fs3_4_PyrLvls2.py:133-139, verbatim call shape
    alphas = [21, 45]
    hsAdapter = HSOpticalFlowAlgoAdapter(alphas, 600)
    lsAdapter = LiuShenOpticalFlowAlgoAdapter(5)
    [U, V] = genericPyramidalOpticalFlow(Iold, Inew, 3.4, hsAdapter, 2, 1, 0.48, lsAdapter)
    assert alphas == [] and U.dtype == np.float32 and U.shape == (512, 512)
    close(U, configs_bundled["c3_U"], TOL_FLOW)
    close(V, configs_bundled["c3_V"], TOL_FLOW)
    with pytest.raises(IndexError):           # second call on the same adapter: alphas exhausted
        genericPyramidalOpticalFlow(Iold, Inew, 3.4, hsAdapter, 2, 1, 0.48, lsAdapter)
    # examples/PyHSchunck_Fs3_4.py:138 passes lsAdapter=None positionally into the FILTER_OPT slot
    [U, V] = genericPyramidalOpticalFlow(Iold, Inew, 3.4, HSOpticalFlowAlgoAdapter([21], 600), 1, 1, None)
    close(U, configs_bundled["c1_U"], TOL_FLOW)
    # benchmark_of_methods.py:156-174
    s = configs_small
    w = GenericPyramidalOpticalFlowWrapper(LiuShenOpticalFlowAlgoAdapter(0.1), filter_sigma=3.4, pyr_levels=2)
    U, V = w.calculateFlow(s["crop0"], s["crop1"])
    close(U, s["bom_LiuSE_HS_Fs3_4_PyrLvls2_U"], TOL_FLOW)
    close(V, s["bom_LiuSE_HS_Fs3_4_PyrLvls2_V"], TOL_FLOW)
    w = GenericPyramidalOpticalFlowWrapper(HSOpticalFlowAlgoAdapter([1.0, 1.0], 100), filter_sigma=3.4, pyr_levels=2)
    Ub, Vb = w.calculateFlowBatch(np.stack([s["crop0"]] * 2), np.stack([s["crop1"]] * 2))
    close(Ub[1], s["bom_HS_Fs3_4_PyrLvls2_U"], TOL_FLOW)
    # adapters stand-alone (the plugin protocol) and the in-place Gaussian
    hs = HSOpticalFlowAlgoAdapter([3.0], 7)
    st = np.load(os.path.join(os.path.dirname(__file__), "golden", "stages.npz"))
    U, V, err = hs.compute(st["hs_f1"], st["hs_f2"], st["hs_U0"], st["hs_V0"])
    close(U, st["hs_U_7"], 2e-6)
    img = st["gauss_in"].copy()
    out = GF.gaussian_filterPx(img, 3.4, 3)
    assert out is img
    same(img, st["gauss_out_g34_3"])


def test_dropin_liu_shen_warp(ofri, lswarp):
    """biLinear=False through the drop-in modules: native adapters (one call) and a foreign adapter (generic path, where
    the warp mutates the level frame in place exactly like the reference)."""
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow
        from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter
    finally:
        sys.path.remove(ofri.SRC_DIR)
    g = lswarp
    U, V = genericPyramidalOpticalFlow(g["crop0"].copy(), g["crop1"].copy(), 3.4, LiuShenOpticalFlowAlgoAdapter(0.1), 2, 1,
                                       biLinear=False)
    close(U, g["lsmain_U"], TOL_FLOW)
    close(V, g["lsmain_V"], TOL_FLOW)

    class ForeignLS(object):                     # the ORACLE's Liu-Shen wrapped as a third-party plugin
        def compute(self, im1, im2, U, V):
            Un, Vn, err, _ = O.ls_compute(im1, im2, 0.1, U, V)
            return Un, Vn, err

        def getAlgoName(self):
            return "foreign LS"

        def hasGenericPyramidalDefaults(self):
            return False

    U, V = genericPyramidalOpticalFlow(g["crop0"].copy(), g["crop1"].copy(), 3.4, ForeignLS(), 2, 1, biLinear=False)
    same(U, g["lsmain_U"])                       # GPU stages bit-exact + oracle LS bit-exact
    same(V, g["lsmain_V"])


def test_dropin_foreign_adapter_generic_path(ofri, configs_small):
    """A third-party duck-typed adapter (here: the ORACLE's HS wrapped as a foreign plugin) drives the generic path:
    GPU stages + adapter.compute on numpy arrays."""
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow
    finally:
        sys.path.remove(ofri.SRC_DIR)

    class Foreign(object):
        def __init__(self):
            self.alphas = [1.0, 1.0]

        def compute(self, im1, im2, U, V):
            return O.hs_compute(im1, im2, self.alphas.pop(), 100, U, V)

        def getAlgoName(self):
            return "foreign HS"

        def hasGenericPyramidalDefaults(self):
            return True

        def getGenericPyramidalDefaults(self):
            return {"warping": True, "biLinear": True, "scaling": True}

    s = configs_small
    U, V = genericPyramidalOpticalFlow(s["crop0"], s["crop1"], 3.4, Foreign(), 2, 1)
    same(U, s["bom_HS_Fs3_4_PyrLvls2_U"])      # GPU stages are bit-exact and the oracle HS is bit-exact
    same(V, s["bom_HS_Fs3_4_PyrLvls2_V"])


def test_dropin_foreign_main_native_liu_shen_refinement(ofri, configs_bundled, bundled_pair):
    """SURVEY 8f-3 (examples/LiuSE_denseLK_Fs2_0_PyrLvls2.py:68-74, LiuSE_Farneback_Fs0_0_PyrLvls2.py:70-76): a FOREIGN
    main adapter refined by this package's Liu-Shen adapter.  Foreign main = the oracle's Horn-Schunck wrapped as a
    third-party plugin, optional = the native Liu-Shen adapter, BASELINE config 3 parameters on the bundled pair: the
    level stages, the refinement and the accumulation stay on the GPU (ofri_pyramidal_flow_external), only the foreign
    compute() runs on the host -- and the result must match the reference's c3 golden."""
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow
        from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter
    finally:
        sys.path.remove(ofri.SRC_DIR)
    calls = []

    class ForeignHS(object):
        def __init__(self):
            self.alphas = [21, 45]

        def compute(self, im1, im2, U, V):
            calls.append(im1.shape)
            return O.hs_compute(im1, im2, self.alphas.pop(), 600, U, V)

        def getAlgoName(self):
            return "foreign HS"

        def hasGenericPyramidalDefaults(self):
            return True

        def getGenericPyramidalDefaults(self):
            return {"warping": True, "biLinear": True, "scaling": True}

    I0, I1 = bundled_pair
    h = ofri.default_handle(0)
    n0 = h.launch_count
    U, V = genericPyramidalOpticalFlow(I0, I1, 3.4, ForeignHS(), 2, 1, 0.48, LiuShenOpticalFlowAlgoAdapter(5))
    assert calls == [(256, 256), (512, 512)]
    assert h.get_option("last_ls_fuse") > 0 and h.launch_count - n0 > 30        # the refinement ran natively
    Ur, Vr = configs_bundled["c3_U"], configs_bundled["c3_V"]
    du, dv = float(np.max(np.abs(U - Ur))), float(np.max(np.abs(V - Vr)))
    Ut, Vt = O.poiseuille_truth(512, 512)
    de = abs(O.epe_rmse(U, V, Ut, Vt) - O.epe_rmse(Ur, Vr, Ut, Vt))
    print("\nforeign HS + native Liu-Shen vs reference c3: max|dU| %.3g max|dV| %.3g |dEPE-RMSE| %.3g" % (du, dv, de))
    assert du <= TOL_FLOW and dv <= TOL_FLOW and de <= TOL_EPE, (du, dv, de)

    class Failing(ForeignHS):
        def compute(self, im1, im2, U, V):
            raise RuntimeError("adapter exploded")

    with pytest.raises(RuntimeError, match="adapter exploded"):      # the adapter's own exception reaches the caller
        genericPyramidalOpticalFlow(I0, I1, 3.4, Failing(), 2, 1, 0.48, LiuShenOpticalFlowAlgoAdapter(5))


def test_sequence_pipeline(ofri, tmp_path):
    """File -> GPU -> file pipeline (opticalflow_ri_b200/pipeline.py): packbits TIFF frames decoded into a
    ring of page-locked slots, one native call per slot, .mat files in the reference's layout -- results identical to the
    direct call on the same frames."""
    Image = pytest.importorskip("PIL.Image")
    sio = pytest.importorskip("scipy.io")
    from opticalflow_ri_b200.pipeline import SequencePipeline
    H, W, n = 96, 160, 7
    frames = []
    for i in range(n + 1):
        a, b = O.synthetic_piv_pair(H, W, seed=40 + i // 2)
        f = a if i % 2 == 0 else b
        frames.append(f)
        Image.fromarray(f.astype(np.uint8)).save(tmp_path / ("fr_%02d.tif" % i), compression="packbits")
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 40), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    hd = ofri.Handle(0)
    try:
        pin = hd.pinned_empty((3, 5))
        pin[...] = 7.0
        assert pin.shape == (3, 5) and float(pin.sum()) == 105.0
        pipe = SequencePipeline(hd, mk, H, W, batch=3, ring=2, decode_workers=2)
        pairs = [(str(tmp_path / ("fr_%02d.tif" % i)), str(tmp_path / ("fr_%02d.tif" % (i + 1))), "flow_%02d" % i)
                 for i in range(n)]
        out = tmp_path / "out"
        st = pipe.run(pairs, out_dir=str(out))
        assert st["pairs"] == n and hd.get_option("last_host_path") == 1          # pinned slots: direct asynchronous copies
        Ud, Vd = hd.pyramidal_flow(np.stack(frames[:-1]), np.stack(frames[1:]), mk())
        for i in range(n):
            m = sio.loadmat(str(out / ("flow_%02d.mat" % i)), squeeze_me=True, struct_as_record=False)
            same(m["velocities"].u, Ud[i])
            same(m["velocities"].v, Vd[i])
    finally:
        hd.close()


def test_pageable_bounce_ring_equals_direct(ofri):
    """Host-pointer call with PAGEABLE numpy arrays over several chunks: the pinned bounce ring (two host threads) must
    deliver exactly what the direct-copy path delivers, errors included."""
    rng = np.random.default_rng(5)
    a = np.stack([rand_img(rng, 64, 80) for _ in range(11)])
    b = np.stack([rand_img(rng, 64, 80) for _ in range(11)])
    mk = lambda: ofri.make_params(ofri.hs_algo([45, 21], 20), ofri.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                  pyramid_levels=2, **HS_DEF)
    hd = ofri.Handle(0)
    try:
        hd.set_option("chunk_pairs", 3)                      # 4 chunks, the last one ragged
        U, V, E = hd.pyramidal_flow(a, b, mk(), want_errors=True)
        assert hd.get_option("last_host_path") == 2
        hd.set_option("host_bounce", 0)
        U0, V0, E0 = hd.pyramidal_flow(a, b, mk(), want_errors=True)
        assert hd.get_option("last_host_path") == 1
        same(U, U0)
        same(V, V0)
        same(E, E0)
    finally:
        hd.close()
