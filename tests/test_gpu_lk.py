"""GPU parity tests of the dense Lucas-Kanade adapter (csrc/ofri_lk.cu behind
src/denseLucasKanade_PyCL.py) against oracle/ofri_lk_oracle.c on seeded inputs.  Oracle and kernel make the same three
choices where the OpenCL original defers to the device (full-float32 bilinear sampler, fused mad, IEEE division) and add
in the same order, so the stand-alone compute() is required to be BIT-IDENTICAL to the oracle."""
import os
import sys

import numpy as np
import pytest

import ofri_lk_oracle as LKO
import ofri_oracle as O
from test_farneback_cpu import piv_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ofri():
    import opticalflow_ri_b200 as o
    return o


@pytest.fixture(scope="module")
def h(ofri):
    hd = ofri.Handle(0)
    yield hd
    hd.close()


@pytest.fixture(scope="module")
def mods(ofri):
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import denseLucasKanade_PyCL as LK
        import GenericPyramidalOpticalFlow as G
        import PhysicsBasedOpticalFlowLiuShen as LS
    finally:
        sys.path.remove(ofri.SRC_DIR)
    return LK, G, LS


CASES = [dict(n_iters=5, half_window=13, asym=(0, 0, 0, 0), init=False),
         dict(n_iters=5, half_window=13, asym=(0, 1, 0, 1), init=True),
         dict(n_iters=3, half_window=5, asym=(0, 0, 0, 0), init=True),      # window < 16: the kernel's other weight rule
         dict(n_iters=8, half_window=7, asym=(1, 0, 0, 1), init=False),     # 15 x 15
         dict(n_iters=5, half_window=15, asym=(0, 0, 0, 0), init=True),     # 31 x 31, the largest the sample grid holds
         dict(n_iters=0, half_window=13, asym=(0, 0, 0, 0), init=True)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_compute_bit_identical_to_oracle(ofri, h, case):
    c = CASES[case]
    a, b = piv_pair(60 + case, 75, 90, shift=(1.1, -0.6))
    a[:20, :30] = 3.0                      # a flat corner: singular structure tensor, flow must pass through untouched
    rng = np.random.default_rng(case)
    U0 = rng.uniform(-0.8, 0.8, a.shape).astype(np.float32) if c["init"] else np.zeros_like(a)
    V0 = rng.uniform(-0.8, 0.8, a.shape).astype(np.float32) if c["init"] else np.zeros_like(a)
    Uo, Vo = LKO.lk_compute(a, b, U0, V0, c["n_iters"], c["half_window"], c["asym"])
    U, V = h.lk_compute(a, b, U0, V0, ofri.lk_params(c["n_iters"], c["half_window"], c["asym"]))
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("lucas-kanade case %d: max|d| %.3g, bit-identical px %.4f" % (case, d, np.mean((U == Uo) & (V == Vo))))
    assert np.array_equal(U, Uo) and np.array_equal(V, Vo)
    assert np.array_equal(U[:3, :10], U0[:3, :10])      # windows (incl. the derivative's reach) entirely inside the flat corner


def test_compute_batched_equals_single(ofri, h, mods):
    LK = mods[0]
    lk = LK.denseLucasKanade_PyCl(Niter=4, halfWindow=9)
    pairs = [piv_pair(80 + i, 64, 70) for i in range(3)]
    A = np.stack([p[0] for p in pairs])
    B = np.stack([p[1] for p in pairs])
    U, V = h.lk_compute(A, B, None, None, lk.native_params())
    for i in range(3):
        u, v, e = lk.compute(A[i], B[i], np.zeros_like(A[i]), np.zeros_like(A[i]))
        assert e is True and np.array_equal(U[i], u) and np.array_equal(V[i], v)


def test_driver_with_lucas_kanade_main(mods):
    """The reference's dense-LK examples (examples/denseLK_Fs2_0.py, LiuSE_denseLK_Fs2_0_PyrLvls2.py:68-74): LK as the main
    adapter (its defaults: no warping, intermediate scaling, no final scaling), alone and refined by Liu-Shen."""
    LK, G, LS = mods
    a, b = piv_pair(9, 112, 128, shift=(1.6, -1.1))
    Uo, Vo = O.pyramidal_flow(a, b, 2.0, LKO.LKParams(5, 13), 2, 1)[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, 2.0, LK.denseLucasKanade_PyCl(Niter=5, halfWindow=13), 2, 1)
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("driver LK: max|d| %.3g" % d)
    assert d <= 1e-4
    Uo, Vo = O.pyramidal_flow(a, b, 2.0, LKO.LKParams(5, 13), 2, 1, 0.48, O.LSParams(4.0))[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, 2.0, LK.denseLucasKanade_PyCl(Niter=5, halfWindow=13), 2, 1, 0.48,
                                         LS.LiuShenOpticalFlowAlgoAdapter(4.0))
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("driver LK + LS: max|d| %.3g" % d)
    assert d <= 1e-4
    A = np.stack([a, b])
    B = np.stack([b, a])
    Ub, Vb = G.genericPyramidalOpticalFlowBatch(A, B, 2.0, LK.denseLucasKanade_PyCl(Niter=5, halfWindow=13), 2, 1, 0.48,
                                                LS.LiuShenOpticalFlowAlgoAdapter(4.0))
    assert np.array_equal(Ub[0], U) and np.array_equal(Vb[0], V)


def test_driver_with_vorticity_enhancement(mods):
    """enableVorticityEnhancement: the window switches depend on the flow of every call, so the driver hands each level
    to the adapter's compute() (host callback, the kernel itself still on the GPU)."""
    LK, G, LS = mods
    yy, xx = np.mgrid[0:96, 0:104].astype(np.float64)
    a, _ = piv_pair(12, 96, 104)
    # second frame = first frame rotated by a small angle about the centre (vorticity of one sign)
    from scipy.ndimage import map_coordinates
    th = 0.02
    xs = 52 + (xx - 52) * np.cos(th) - (yy - 48) * np.sin(th)
    ys = 48 + (xx - 52) * np.sin(th) + (yy - 48) * np.cos(th)
    b = map_coordinates(a.astype(np.float64), [ys, xs], order=3, mode='nearest').astype(np.float32)
    Uo, Vo = O.pyramidal_flow(a, b, 2.0, LKO.LKParams(5, 13, enableVorticityEnhancement=True), 2, 1)[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, 2.0, LK.denseLucasKanade_PyCl(enableVorticityEnhancement=True), 2, 1)
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("driver LK + vorticity switches: max|d| %.3g" % d)
    assert d <= 1e-4
    # the switches were really on at the second level
    Un, Vn = O.pyramidal_flow(a, b, 2.0, LKO.LKParams(5, 13), 2, 1)[:2]
    assert not np.array_equal(Un, Uo)


def test_lk_argument_errors(ofri, h):
    z = np.zeros((40, 40), np.float32)
    p = ofri.lk_params()
    p.size = 8
    with pytest.raises(ValueError, match="ABI"):
        h.lk_compute(z, z, None, None, p)
    p = ofri.lk_params(asym=(0, 2, 0, 0))
    with pytest.raises(ValueError, match="switches"):
        h.lk_compute(z, z, None, None, p)
    h2 = ofri.Handle(0)
    try:
        params = ofri.make_params(ofri.lk_algo(), None, filter_sigma=0.0, pyramid_levels=1, k_levels=1)
        with pytest.raises(ValueError, match="ofri_set_lk"):
            h2.pyramidal_flow(z, z, params)
    finally:
        h2.close()


EXAMPLES = {   # the reference's example scripts (examples/*.py) beyond the three BASELINE ones: (main, FILTER, levels, LS h, kwargs)
    "Farneback_Fs0_0": ("fb", 0, 1, None, {}),
    "Farneback_Fs0_0_PyrLvls2": ("fb", 0, 2, None, {}),
    "LiuSE_Farneback_Fs0_0_PyrLvls2": ("fb", 0, 2, 10, {}),
    "denseLK_Fs2_0": ("lk", 2, 1, None, dict(warping=False)),
    "LiuSE_denseLK_Fs2_0_PyrLvls2": ("lk", 2, 2, 10, dict(warping=False)),
}


@pytest.mark.parametrize("name", sorted(EXAMPLES))
def test_reference_example_configurations(ofri, mods, bundled_pair, name):
    """The Farneback / dense-LK example scripts of the reference, with their parameters, on a crop of the bundled
    Poiseuille pair: drop-in modules (GPU) against the oracle driver with the oracle adapters."""
    import ofri_farneback_oracle as FBO
    LK, G, LS = mods
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import Farneback_PyCL as FB
    finally:
        sys.path.remove(ofri.SRC_DIR)
    main, FILTER, levels, ls_h, kw = EXAMPLES[name]
    a = np.ascontiguousarray(bundled_pair[0][180:340, 160:336])
    b = np.ascontiguousarray(bundled_pair[1][180:340, 160:336])
    if main == "fb":
        ours, theirs = FB.Farneback_PyCL(platformID=0), FBO.FBParams()
    else:
        ours, theirs = LK.denseLucasKanade_PyCl(Niter=5, halfWindow=13, platformID=0), LKO.LKParams(5, 13)
    Uo, Vo = O.pyramidal_flow(a, b, FILTER, theirs, levels, 1, 0.48, O.LSParams(ls_h) if ls_h else None, **kw)[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, FILTER, ours, levels, 1, 0.48,
                                         LS.LiuShenOpticalFlowAlgoAdapter(ls_h) if ls_h else None, **kw)
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("%s: max|d| %.3g px, flow range U [%.3f, %.3f]" % (name, d, U.min(), U.max()))
    assert d <= 1e-4


@pytest.mark.parametrize("shape", [(8, 8), (19, 23), (33, 50), (5, 131)])
def test_small_and_odd_shapes(ofri, h, shape):
    """Frames smaller than the window, widths that are not a multiple of 4 (row pitch != W on the device), batch of 2."""
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    H, W = shape
    A = rng.uniform(0, 255, (2, H, W)).astype(np.float32)
    B = np.roll(A, 1, 2) + rng.uniform(-2, 2, (2, H, W)).astype(np.float32)
    U0 = rng.uniform(-0.5, 0.5, (2, H, W)).astype(np.float32)
    U, V = h.lk_compute(A, B, U0, -U0, ofri.lk_params(5, 13))
    for i in range(2):
        uo, vo = LKO.lk_compute(A[i], B[i], U0[i], -U0[i], 5, 13)
        assert np.array_equal(U[i], uo) and np.array_equal(V[i], vo), (shape, i)


def test_farneback_and_lk_as_optional_adapters(ofri, mods):
    """The OpenCL adapters in the OPTIONAL slot (refining a Horn-Schunck main adapter), and both together (dense LK main,
    Farneback optional): kinds OFRI_ALGO_FB / OFRI_ALGO_LK in ofri_params.opt_algo."""
    import ofri_farneback_oracle as FBO
    LK, G, LS = mods
    sys.path.insert(0, ofri.SRC_DIR)
    try:
        import Farneback_PyCL as FB
        import HornSchunck as HS
    finally:
        sys.path.remove(ofri.SRC_DIR)
    a, b = piv_pair(31, 96, 112, shift=(1.2, 0.7))
    fkw = dict(windowSize=13, Niters=2, polyN=5, polySigma=1.1)
    cases = [(lambda: HS.HSOpticalFlowAlgoAdapter([8.0, 12.0], 40), lambda: O.HSParams([8.0, 12.0], 40),
              lambda: FB.Farneback_PyCL(**fkw), lambda: FBO.FBParams(**fkw)),
             (lambda: HS.HSOpticalFlowAlgoAdapter([8.0, 12.0], 40), lambda: O.HSParams([8.0, 12.0], 40),
              lambda: LK.denseLucasKanade_PyCl(Niter=3, halfWindow=7), lambda: LKO.LKParams(3, 7)),
             (lambda: LK.denseLucasKanade_PyCl(Niter=3, halfWindow=7), lambda: LKO.LKParams(3, 7),
              lambda: FB.Farneback_PyCL(**fkw), lambda: FBO.FBParams(**fkw))]
    for i, (m, mo, o, oo) in enumerate(cases):
        Uo, Vo = O.pyramidal_flow(a, b, 2.0, mo(), 2, 1, 0.48, oo())[:2]
        U, V = G.genericPyramidalOpticalFlow(a, b, 2.0, m(), 2, 1, 0.48, o())
        d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
        print("optional-slot case %d: max|d| %.3g" % (i, d))
        assert d <= 1e-4, (i, d)


def test_benchmark_of_methods_table(ofri, bundled_pair):
    """examples/run_benchmark_of_methods.py: the ten rows of the reference's benchmark_of_methods.py; its six dense-LK /
    Farneback rows against the oracle driver on a crop of the bundled pair (the four HS rows have reference goldens:
    test_gpu_parity.py::test_driver_bom_rows)."""
    import importlib.util
    import ofri_farneback_oracle as FBO
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bom_b200", os.path.join(root, "examples", "run_benchmark_of_methods.py"))
    bom = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bom)
    a = np.ascontiguousarray(bundled_pair[0][200:328, 180:324])
    b = np.ascontiguousarray(bundled_pair[1][200:328, 180:324])
    rows = [r for r in bom.ROWS if r[1] in ("lk", "fb")]
    res = bom.run_benchmark(a, b, None, rows)
    for name, family, sigma, levels, ls in rows:
        theirs = O.LSParams(0.1) if ls else (LKO.LKParams(5, 13) if family == "lk" else FBO.FBParams())
        Uo, Vo = O.pyramidal_flow(a, b, sigma, theirs, levels, 1)[:2]
        d = max(np.abs(res[name]["U"] - Uo).max(), np.abs(res[name]["V"] - Vo).max())
        print("BOM row %s: max|d| %.3g" % (name, d))
        assert d <= 1e-4, (name, d)
