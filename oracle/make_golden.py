#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/src).

Run in the build container only (`python oracle/make_golden.py`); the GPU box has no
/root/reference, so the vectors are committed.  The reference has no tests or golden vectors of
its own -- these files are what pins the oracle (oracle/ofri_oracle.py) and, through
it, the CUDA path.  matplotlib is not installed and is never called on the path, so it is stubbed
(SURVEY §A.9); TIFFs are read with PIL (identical uint8 arrays to skimage.io.imread).
"""
import builtins
import contextlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    mpl, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    for n in ("imshow", "figure", "show"):
        setattr(pp, n, lambda *a, **k: None)
    mpl.pyplot = pp
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = pp
    sys.path.insert(0, os.path.join(REF, "src"))
    import warnings
    warnings.simplefilter("ignore")
    import GenericPyramidalOpticalFlow as GPOF
    import HornSchunck as HS
    import PhysicsBasedOpticalFlowLiuShen as LS
    import gaussian_filter as GF
    import GaussianKernelBitExact as GKBE
    import GenericPyramidalOpticalFlowWrapper as WRAP
    return GPOF, HS, LS, GF, GKBE, WRAP


@contextlib.contextmanager
def quiet():
    saved = builtins.print
    builtins.print = lambda *a, **k: None
    try:
        yield
    finally:
        builtins.print = saved


def load_bundled():
    from PIL import Image
    base = os.path.join(REF, "examples", "testImages", "Bits08", "Ni06")
    a = np.array(Image.open(os.path.join(base, "parabolic01_0.tif")))
    b = np.array(Image.open(os.path.join(base, "parabolic01_1.tif")))
    assert a.dtype == np.uint8 and a.shape == (512, 512)
    return a, b


def rand_img(rng, H, W):
    """PIV-like random test image: sparse bright blobs on a dark background, values 0..255."""
    img = rng.uniform(0, 12, (H, W))
    n = max(4, H * W // 40)
    ys = rng.integers(0, H, n)
    xs = rng.integers(0, W, n)
    img[ys, xs] += rng.uniform(40, 240, n)
    return np.clip(np.rint(img), 0, 255).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    GPOF, HS, LS, GF, GKBE, WRAP = import_reference()
    rng = np.random.default_rng(20261018)
    a8, b8 = load_bundled()
    np.savez_compressed(os.path.join(OUT, "bundled_pair.npz"), im0=a8, im1=b8)
    I0 = a8.astype(np.float32)
    I1 = b8.astype(np.float32)

    st = {}
    # ---- Gaussian coefficient + filter (GF) ------------------------------------------------
    for tag, (sg, K) in {"g34_3": (3.4, 3), "g048_5": (0.48, 5), "g20_3": (2.0, 3), "g12_7": (1.2, 7),
                         "g18_9": (1.8, 9)}.items():
        st["gk_" + tag] = GF.prepareGaussianKernel(sg, K)
    gimg = rand_img(rng, 37, 53)
    st["gauss_in"] = gimg
    for tag, (sg, K) in {"g34_3": (3.4, 3), "g048_5": (0.48, 5), "g12_7": (1.2, 7), "g18_9": (1.8, 9)}.items():
        st["gauss_out_" + tag] = GF.gaussian_filterPx(np.copy(gimg), sg, K)
    st["gauss_trunc_out"] = GF.gaussian_filter(np.copy(gimg), 1.8, 4.0 / 0.6 * 3)     # GPOF:211 parameters
    gsmall = rand_img(rng, 5, 6)
    st["gauss_small_in"] = gsmall
    st["gauss_small_out"] = GF.gaussian_filterPx(np.copy(gsmall), 0.48, 5)
    # ---- Pillow bicubic down-sample (GPOF:67-68) --------------------------------------------
    for tag, (H, W) in {"a": (48, 64), "b": (47, 61), "c": (33, 50), "d": (150, 259)}.items():
        img = rand_img(rng, H, W)
        w = int(np.int32(np.round(W * 0.5)))
        h = int(np.int32(np.round(H * 0.5)))
        st["rs_in_" + tag] = img
        st["rs_out_" + tag] = GPOF.imresize(img, (w, h))
    img = rand_img(rng, 64, 96)
    st["rs_in_q"] = img
    st["rs_out_q"] = GPOF.imresize(img, (24, 16))          # 4:1, as level 1 of a 3-level pyramid
    # ---- spline up-sample + warp through updateNextPyramidalLevel (GPOF:118-235) -----------
    for tag, (h, w, H, W) in {"a": (24, 32, 48, 64), "b": (24, 30, 47, 61), "c": (16, 25, 33, 50),
                              "d": (16, 24, 64, 96)}.items():
        yy, xx = np.mgrid[0:h, 0:w]
        Ua = (1.5 * np.sin(yy / 5.0) + 0.8 * np.cos(xx / 7.0) + rng.normal(0, 0.2, (h, w))).astype(np.float32)
        Va = (0.9 * np.cos(yy / 4.0) * np.sin(xx / 6.0) + rng.normal(0, 0.2, (h, w))).astype(np.float32)
        n1 = rand_img(rng, H, W)
        n2 = rand_img(rng, H, W)
        prev = np.zeros((h, w), dtype=np.float32)
        st["up_Ua_" + tag] = Ua.copy()
        st["up_Va_" + tag] = Va.copy()
        st["up_n1_" + tag] = n1.copy()
        st["up_n2_" + tag] = n2.copy()
        for sc in (False, True):
            with quiet():
                w1, w2, Uacc, Vacc, U, V = GPOF.updateNextPyramidalLevel(n1.copy(), prev, n2.copy(), Ua.copy(),
                                                                         Va.copy(), None, None, True, True, sc)
            s = "_s1_" if sc else "_s0_"
            st["up_w1" + s + tag] = w1
            st["up_w2" + s + tag] = w2
            st["up_Uacc" + s + tag] = Uacc
            st["up_Vacc" + s + tag] = Vacc
    # direct warp with large / out-of-range coordinates (GPOF:70-116)
    img = rand_img(rng, 29, 41)
    cy = (np.arange(29, dtype=np.float32)[:, None] + rng.uniform(-6, 6, (29, 41))).astype(np.float32)
    cx = (np.arange(41, dtype=np.float32)[None, :] + rng.uniform(-6, 6, (29, 41))).astype(np.float32)
    cy[3, 4] = 7.5
    cx[3, 4] = 8.5
    cy[5, 6] = -0.5
    cx[5, 6] = 40.5          # exact .5 ties -> half-to-even
    st["warp_img"] = img
    st["warp_cy"] = cy
    st["warp_cx"] = cx
    st["warp_out"] = GPOF.doBiLinearWarping(img, cy.copy(), cx.copy(), 1, "nearest")
    # ---- Horn-Schunck (HS) ----------------------------------------------------------------
    f1 = GF.gaussian_filterPx(rand_img(rng, 45, 58), 3.4, 3)
    f2 = GF.gaussian_filterPx(rand_img(rng, 45, 58), 3.4, 3)
    st["hs_f1"] = f1
    st["hs_f2"] = f2
    fx, fy, ft = HS.computeDerivatives(f2, f1)        # as reached from compute(f1, f2): HS:37, 73, 84
    st["hs_fx"], st["hs_fy"], st["hs_ft"] = fx, fy, ft
    U0 = rng.normal(0, 0.5, (45, 58)).astype(np.float32)
    V0 = rng.normal(0, 0.5, (45, 58)).astype(np.float32)
    st["hs_U0"], st["hs_V0"] = U0, V0
    for nit in (1, 2, 7, 50):
        with quiet():
            U, V, err = HS.HSOpticalFlowAlgoAdapter([3.0], nit).compute(f1, f2, U0.copy(), V0.copy())
        st["hs_U_%d" % nit], st["hs_V_%d" % nit], st["hs_err_%d" % nit] = U, V, np.float64(err)
    # ---- Liu-Shen (LS) ---------------------------------------------------------------------
    g1 = GF.gaussian_filterPx(rand_img(rng, 40, 52), 0.48, 5)
    g2 = GF.gaussian_filterPx(rand_img(rng, 40, 52), 0.48, 5)
    st["ls_g1"], st["ls_g2"] = g1, g2
    rec = []
    orig_helper = LS.helper

    def recording_helper(B11, B12, B22, bu, bv, u, v, r, c):
        out = orig_helper(B11, B12, B22, bu, bv, u, v, r, c)
        rec.append((B11.copy(), B12.copy(), B22.copy(), bu.copy(), bv.copy(), out[0].copy(), out[1].copy(),
                    float(out[2])))
        return out

    LS.helper = recording_helper
    Uin = rng.normal(0, 0.3, (40, 52)).astype(np.float32)
    Vin = rng.normal(0, 0.3, (40, 52)).astype(np.float32)
    st["ls_Uin"], st["ls_Vin"] = Uin, Vin
    for tag, hpar in {"h5": 5, "h01": 0.1}.items():
        rec.clear()
        with quiet():
            U, V, err = LS.LiuShenOpticalFlowAlgoAdapter(hpar).compute(g1.copy(), g2.copy(), Uin.copy(), Vin.copy())
        st["ls_U_" + tag], st["ls_V_" + tag], st["ls_err_" + tag] = U, V, np.float64(err)
        st["ls_niter_" + tag] = np.int64(len(rec))
        B11, B12, B22, bu, bv, un, vn, e0 = rec[0]
        st["ls_B11_" + tag], st["ls_B12_" + tag], st["ls_B22_" + tag] = B11, B12, B22
        st["ls_bu0_" + tag], st["ls_bv0_" + tag] = bu, bv
        st["ls_u1_" + tag], st["ls_v1_" + tag], st["ls_err0_" + tag] = un, vn, np.float64(e0)
        st["ls_errs_" + tag] = np.array([r[-1] for r in rec], dtype=np.float64)
    # zero initial guess: iteration-0 bu/bv are exactly Ixt/Iyt
    rec.clear()
    with quiet():
        LS.LiuShenOpticalFlowAlgoAdapter(5).compute(g1.copy(), g2.copy(), np.zeros_like(Uin), np.zeros_like(Vin))
    st["ls_Ixt"], st["ls_Iyt"] = rec[0][3], rec[0][4]
    # early exit: identical frames -> zero flow -> total_error == 0 after the first sweep
    rec.clear()
    with quiet():
        U, V, err = LS.LiuShenOpticalFlowAlgoAdapter(5).compute(g1.copy(), g1.copy(), np.zeros_like(Uin),
                                                                np.zeros_like(Vin))
    st["ls_same_niter"] = np.int64(len(rec))
    st["ls_same_U"], st["ls_same_V"], st["ls_same_err"] = U, V, np.float64(err)
    LS.helper = orig_helper
    # ---- getGaussianKernelBitExact (GKBE) ----------------------------------------------------
    for tag, (n, sg) in {"3_0": (3, 0.0), "5_0": (5, 0.0), "7_0": (7, -0.0), "9_0": (9, 0), "3_34": (3, 3.4),
                         "5_048": (5, 0.48), "7_15": (7, 1.5), "33_495": (33, 4.95), "4_1": (4, 1.0),
                         "5_m12": (5, -1.2), "11_0": (11, 0.0)}.items():
        s, kv = GKBE.getGaussianKernelBitExact(n, sg)
        st["gkbe_sum_" + tag] = np.float64(s)
        st["gkbe_k_" + tag] = np.asarray(kv, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **st)

    # ---- whole-driver goldens ----------------------------------------------------------------
    def run(im0, im1, FILTER, main, L=1, k=1, FILTER_OPT=None, opt=None, **kw):
        with quiet():
            U, V = GPOF.genericPyramidalOpticalFlow(im0, im1, FILTER, main, L, k, FILTER_OPT, opt, **kw)
        return np.asarray(U, dtype=np.float32), np.asarray(V, dtype=np.float32)

    cfg = {}
    U, V = run(I0, I1, 3.4, HS.HSOpticalFlowAlgoAdapter([21], 600), 1, 1)
    cfg["c1_U"], cfg["c1_V"] = U, V
    U, V = run(I0, I1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 600), 2, 1)
    cfg["c2_U"], cfg["c2_V"] = U, V
    U, V = run(I0, I1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 600), 2, 1, 0.48, LS.LiuShenOpticalFlowAlgoAdapter(5))
    cfg["c3_U"], cfg["c3_V"] = U, V
    np.savez_compressed(os.path.join(OUT, "configs_bundled.npz"), **cfg)

    # smaller whole-driver cases: BOM rows (BOM:143-194) on a 200x232 crop, odd sizes, 3 levels, kLevels=2,
    # no-warp, LS as main, FILTER=0
    sm = {}
    c0 = np.ascontiguousarray(I0[150:350, 100:332])
    c1 = np.ascontiguousarray(I1[150:350, 100:332])
    sm["crop0"], sm["crop1"] = c0, c1
    for name, fs, L, use_ls in (("HS_Fs0_0", 0.0, 1, False), ("HS_Fs3_4", 3.4, 1, False),
                                ("HS_Fs3_4_PyrLvls2", 3.4, 2, False), ("LiuSE_HS_Fs3_4_PyrLvls2", 3.4, 2, True)):
        algo = LS.LiuShenOpticalFlowAlgoAdapter(0.1) if use_ls else HS.HSOpticalFlowAlgoAdapter([1.0] * L, 100)
        with quiet():
            U, V = WRAP.GenericPyramidalOpticalFlowWrapper(algo, filter_sigma=fs, pyr_levels=L).calculateFlow(c0, c1)
        sm["bom_%s_U" % name], sm["bom_%s_V" % name] = np.float32(U), np.float32(V)
    o0 = np.ascontiguousarray(I0[7:7 + 151, 11:11 + 259])          # odd sizes: 151 x 259 -> 76 x 130 (half-even)
    o1 = np.ascontiguousarray(I1[7:7 + 151, 11:11 + 259])
    sm["odd0"], sm["odd1"] = o0, o1
    U, V = run(o0, o1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 200), 2, 1, 0.48, LS.LiuShenOpticalFlowAlgoAdapter(5))
    sm["odd_c3_U"], sm["odd_c3_V"] = U, V
    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 30, 45], 150), 3, 1, 0.48, LS.LiuShenOpticalFlowAlgoAdapter(5))
    sm["l3_U"], sm["l3_V"] = U, V
    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 21, 45, 45], 100), 2, 2, 0.48,
               LS.LiuShenOpticalFlowAlgoAdapter(5))
    sm["k2_U"], sm["k2_V"] = U, V
    U, V = run(c0, c1, 0.8, HS.HSOpticalFlowAlgoAdapter([21, 21, 45, 45], 100), 2, 2)       # k>0 with FILTER<=1
    sm["k2f08_U"], sm["k2f08_V"] = U, V
    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 100, False), 2, 1, warping=False,
               pyramidalIntermediateScaling=True, pyramidalScaling=True)
    sm["nowarp_U"], sm["nowarp_V"] = U, V
    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 21, 45, 45], 60, False), 2, 2, warping=False,
               pyramidalScaling=True)
    sm["nowarp_k2_U"], sm["nowarp_k2_V"] = U, V
    U, V = run(c0, c1, 3.4, LS.LiuShenOpticalFlowAlgoAdapter(5), 2, 1, 0.0, HS.HSOpticalFlowAlgoAdapter([30, 30], 40))
    sm["lsmain_hsopt_U"], sm["lsmain_hsopt_V"] = U, V          # LS main, HS optional, FILTER_OPT = 0 (GPOF:384-386)
    np.savez_compressed(os.path.join(OUT, "configs_small.npz"), **sm)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__" and not ({"--lswarp", "--big", "--extra", "--farneback"} & set(sys.argv)):
    main()


def main_lswarp():
    """`python oracle/make_golden.py --lswarp`: goldens of the biLinear=False "Liu-Shen warp" branch (GPOF:190-221),
    written to tests/golden/configs_lswarp.npz (kept separate so the other files need not be regenerated)."""
    GPOF, HS, LS, GF, GKBE, WRAP = import_reference()
    a8, b8 = load_bundled()
    I0, I1 = a8.astype(np.float32), b8.astype(np.float32)
    c0 = np.ascontiguousarray(I0[150:350, 100:332])
    c1 = np.ascontiguousarray(I1[150:350, 100:332])

    def run(im0, im1, FILTER, main, L=1, k=1, FILTER_OPT=None, opt=None, **kw):
        a, b = im0.copy(), im1.copy()                 # the branch mutates its input frame in place
        with quiet():
            U, V = GPOF.genericPyramidalOpticalFlow(a, b, FILTER, main, L, k, FILTER_OPT, opt, **kw)
        return np.asarray(U, dtype=np.float32), np.asarray(V, dtype=np.float32), a

    g = {"crop0": c0, "crop1": c1}
    # stage level: one transition with a smooth flow field (integer shifts, wrap-around at the left border, collisions)
    rng = np.random.default_rng(9)
    yy, xx = np.mgrid[0:100, 0:116].astype(np.float32)
    # targets must stay inside the frame at the right / bottom border (the reference raises IndexError otherwise) but may
    # be negative at the left / top border (numpy wraps them around)
    Ua = (-3.7 * (1 - ((yy - 49.5) / 50) ** 2) - 0.4 + 0.3 * np.sin(xx / 9)).astype(np.float32)
    Va = (0.8 * (1 - 2 * yy / 99) * np.cos(xx / 13) ** 2 - 0.9 * (yy < 3)).astype(np.float32)
    n1 = c0.copy()
    with quiet():
        w1, w2, ua, va, u0, v0 = GPOF.updateNextPyramidalLevel(n1, np.zeros((100, 116), np.float32), c1.copy(), Ua.copy(),
                                                               Va.copy(), None, None, True, False, True)
    g["st_Ua"], g["st_Va"], g["st_w1"], g["st_w2"], g["st_ua"], g["st_va"] = Ua, Va, np.float32(w1), np.float32(w2), ua, va
    U, V, _ = run(c0, c1, 3.4, LS.LiuShenOpticalFlowAlgoAdapter(0.1), 2, 1, biLinear=False)
    g["lsmain_U"], g["lsmain_V"] = U, V
    U, V, a = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 100, False), 2, 1, 0.48,
                  LS.LiuShenOpticalFlowAlgoAdapter(5), biLinear=False, pyramidalScaling=True)
    g["hsnodef_U"], g["hsnodef_V"], g["hsnodef_im1_after"] = U, V, a
    U, V, _ = run(c0, c1, 3.4, LS.LiuShenOpticalFlowAlgoAdapter(1.0), 3, 1, biLinear=False)
    g["l3_U"], g["l3_V"] = U, V
    # kLevels = 2 on this pair drives a scatter target past the bottom border: the reference raises IndexError
    try:
        run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([21, 21, 45, 45], 60, False), 2, 2, biLinear=False, pyramidalScaling=True)
        g["k2_raises_index_error"] = np.int32(0)
    except IndexError:
        g["k2_raises_index_error"] = np.int32(1)
    np.savez_compressed(os.path.join(OUT, "configs_lswarp.npz"), **g)
    print("configs_lswarp.npz", os.path.getsize(os.path.join(OUT, "configs_lswarp.npz")))


if __name__ == "__main__" and "--lswarp" in sys.argv:
    main_lswarp()


def main_big(which):
    """`python oracle/make_golden.py --big 1024|2048`: the reference's flow for one seeded synthetic PIV pair at the frame
    sizes BASELINE names (config 4: 1024 x 1024; the row-band path: 2048 x 2048), full EX3 parameters
    (examples/LiuSE_PyHSchunck_Fs3_4_PyrLvls2.py:133-139: FILTER 3.4, HS alphas [21, 45] x 600 sweeps, 2 levels,
    FILTER_OPT 0.48, Liu-Shen h = 5).  Inputs are stored as uint8 (what the generator renders), outputs as float32.
    ~1 min / ~4 min of single-core time; written to tests/golden/big_<size>.npz."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair          # host-side generator, not the compute path
    GPOF, HS, LS, GF, GKBE, WRAP = import_reference()
    n = int(which)
    seed = {1024: 0, 2048: 1}.get(n, n)
    a, b = synthetic_piv_pair(n, n, seed)
    import time
    t0 = time.time()
    with quiet():
        U, V = GPOF.genericPyramidalOpticalFlow(a.copy(), b.copy(), 3.4, HS.HSOpticalFlowAlgoAdapter([21, 45], 600), 2, 1,
                                                0.48, LS.LiuShenOpticalFlowAlgoAdapter(5))
    dt = time.time() - t0
    out = os.path.join(OUT, "big_%d.npz" % n)
    np.savez_compressed(out, im0=a.astype(np.uint8), im1=b.astype(np.uint8), U=np.float32(U), V=np.float32(V),
                        seed=np.int64(seed), ref_seconds=np.float64(dt))
    print(out, os.path.getsize(out), "reference took %.1f s" % dt)


def main_extra():
    """`python oracle/make_golden.py --extra`: additional small whole-driver cases (tests/golden/configs_extra.npz):
    Horn-Schunck only, kLevels = 2 with a weak regularisation (alpha = 1): the k-loop re-warps the level images by the
    UNREFINED Horn-Schunck result of the previous k (GPOF:392-404), the case where a float32 coordinate flip in the
    warp is followed 1:1 by the next solve."""
    GPOF, HS, LS, GF, GKBE, WRAP = import_reference()
    a8, b8 = load_bundled()
    I0, I1 = a8.astype(np.float32), b8.astype(np.float32)
    c0 = np.ascontiguousarray(I0[150:350, 100:332])
    c1 = np.ascontiguousarray(I1[150:350, 100:332])
    g = {"crop0": c0, "crop1": c1}

    def run(im0, im1, FILTER, main, L=1, k=1, FILTER_OPT=None, opt=None, **kw):
        with quiet():
            U, V = GPOF.genericPyramidalOpticalFlow(im0, im1, FILTER, main, L, k, FILTER_OPT, opt, **kw)
        return np.asarray(U, dtype=np.float32), np.asarray(V, dtype=np.float32)

    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([1.0, 1.0, 1.0, 1.0], 100), 2, 2)
    g["k2a1_U"], g["k2a1_V"] = U, V
    U, V = run(c0, c1, 3.4, HS.HSOpticalFlowAlgoAdapter([1.0, 1.0], 100), 1, 2)
    g["l1k2a1_U"], g["l1k2a1_V"] = U, V
    U, V = run(c0, c1, 2.0, HS.HSOpticalFlowAlgoAdapter([0.5, 0.5, 2.0, 2.0, 2.0, 2.0], 60), 3, 2)
    g["l3k2_U"], g["l3k2_V"] = U, V
    out = os.path.join(OUT, "configs_extra.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out))


if __name__ == "__main__" and "--big" in sys.argv:
    main_big(sys.argv[sys.argv.index("--big") + 1])
if __name__ == "__main__" and "--extra" in sys.argv:
    main_extra()


def main_farneback():
    """`python oracle/make_golden.py --farneback`: the HOST-SIDE coefficient tables of the reference's Farneback adapter
    (tests/golden/farneback_tables.npz).  The adapter's kernels need an OpenCL runtime this image does not have, so only
    its pure-Python table generators are run (pyopencl is stubbed; the constructor, which opens an OpenCL context, is
    bypassed with object.__new__): FarnebackPrepareGaussian, setPolynomialExpansionConsts, setGaussianBlurKernel."""
    import_reference()
    for n in ("pyopencl", "pyopencl.array"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["pyopencl"].array = sys.modules["pyopencl.array"]
    import Farneback_PyCL as FB
    g = {}
    for tag, (n, sigma) in {"7_15": (7, 1.5), "5_11": (5, 1.1), "5_0": (5, 0.0), "7_12": (7, 1.2)}.items():
        o = object.__new__(FB.Farneback_PyCL)
        o.polyN, o.polySigma = n, sigma
        o.setPolynomialExpansionConsts()
        g["g_" + tag], g["xg_" + tag], g["xxg_" + tag] = o.matrixG[0], o.matrixXG[0], o.matrixXXG[0]
        g["ig_" + tag], g["igd_" + tag] = o.matrixIG, o.matrixIGD
    for tag, (size, sigma) in {"33": (33, 33 / 2 * 0.3), "13": (13, 13 / 2 * 0.3), "3_0": (3, 0.0), "3_05": (3, 0.5),
                               "7_15": (7, 1.5), "17_35": (17, 3.5)}.items():
        o = object.__new__(FB.Farneback_PyCL)
        o.setGaussianBlurKernel(size, sigma)
        g["blur_" + tag] = o.matrixGKernel[0]
    out = os.path.join(OUT, "farneback_tables.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out))


if __name__ == "__main__" and "--farneback" in sys.argv:
    main_farneback()
