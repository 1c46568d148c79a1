"""GPU parity tests of the Farneback adapter (csrc/ofri_farneback.cu behind src/Farneback_PyCL.py) against
oracle/ofri_farneback_oracle.py on seeded inputs.  The oracle rounds every multiply and add separately in the order the
reference's OpenCL kernels write them and so does the CUDA path (explicit __fmul_rn / __fadd_rn), hence the stand-alone
compute() is required to be BIT-IDENTICAL to the oracle (the division in updateFlow is IEEE in both); inside the generic
driver the north-star tolerance of the driver's own stages applies (1e-4 px)."""
import sys

import numpy as np
import pytest

import ofri_farneback_oracle as FBO
import ofri_oracle as O
from test_farneback_cpu import piv_pair

pytestmark = pytest.mark.gpu

TOL = 0.0          # measured on B200: every case bit-identical to the oracle


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
        import Farneback_PyCL as FB
        import GenericPyramidalOpticalFlow as G
        import PhysicsBasedOpticalFlowLiuShen as LS
    finally:
        sys.path.remove(ofri.SRC_DIR)
    return FB, G, LS


def test_bilinear_resample_bit_exact(h):
    from PIL import Image
    rng = np.random.default_rng(5)
    im = rng.uniform(-3, 3, (37, 45)).astype(np.float32)
    for (w, hh) in ((23, 19), (45, 37), (90, 74), (51, 40), (11, 9)):
        want = np.array(Image.fromarray(im).resize((w, hh), Image.BILINEAR))
        assert np.array_equal(h.resize_bilinear(im, hh, w), want), (w, hh)


CASES = [dict(windowSize=13, Niters=3, polyN=5, polySigma=1.1, useGaussian=False, pyramidalLevels=1),
         dict(windowSize=13, Niters=3, polyN=7, polySigma=1.5, useGaussian=True, pyramidalLevels=1),
         dict(windowSize=33, Niters=5, polyN=7, polySigma=1.5, useGaussian=True, pyramidalLevels=1),
         dict(windowSize=15, Niters=2, polyN=5, polySigma=1.1, useGaussian=True, pyramidalLevels=3),
         dict(windowSize=9, Niters=2, polyN=7, polySigma=1.5, useGaussian=False, pyramidalLevels=2, pyrScale=0.6)]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_compute_matches_oracle(mods, case):
    FB = mods[0]
    kw = CASES[case]
    a, b = piv_pair(20 + case, 141, 170)
    rng = np.random.default_rng(case)
    U0 = rng.uniform(-0.5, 0.5, a.shape).astype(np.float32) if case % 2 else np.zeros_like(a)
    V0 = rng.uniform(-0.5, 0.5, a.shape).astype(np.float32) if case % 2 else np.zeros_like(a)
    Uo, Vo, _ = FBO.FBParams(**kw).compute(a, b, U0, V0)
    U, V, err = FB.Farneback_PyCL(**kw).compute(a, b, U0, V0)
    assert err == 'Unknown' and U.dtype == np.float32 and U.shape == a.shape
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("farneback case %d: max|d| %.3g, bit-identical px %.4f" % (case, d, np.mean((U == Uo) & (V == Vo))))
    assert d <= TOL, d


def test_compute_batched_equals_single(h, mods):
    FB = mods[0]
    fb = FB.Farneback_PyCL(windowSize=13, Niters=2, polyN=5, polySigma=1.1, pyramidalLevels=2)
    pairs = [piv_pair(40 + i, 96, 104) for i in range(3)]
    A = np.stack([p[0] for p in pairs])
    B = np.stack([p[1] for p in pairs])
    U, V = h.farneback_compute(A, B, None, None, fb.native_params())
    for i in range(3):
        u, v, _ = fb.compute(A[i], B[i], np.zeros_like(A[i]), np.zeros_like(A[i]))
        assert np.array_equal(U[i], u) and np.array_equal(V[i], v)


def test_driver_with_farneback_main(mods):
    """Farneback as the main adapter of the generic driver (its defaults: warping False, final scaling True), alone and
    refined by Liu-Shen (the reference's Farneback + Liu-Shen example configuration), against the oracle driver."""
    FB, G, LS = mods
    a, b = piv_pair(7, 128, 144, shift=(1.6, -1.1))
    kw = dict(windowSize=13, Niters=3, polyN=7, polySigma=1.5, pyramidalLevels=1)
    Uo, Vo = O.pyramidal_flow(a, b, 0.0, FBO.FBParams(**kw), 2, 1)[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, 0.0, FB.Farneback_PyCL(**kw), 2, 1)
    assert max(np.abs(U - Uo).max(), np.abs(V - Vo).max()) <= 1e-4
    Uo, Vo = O.pyramidal_flow(a, b, 2.0, FBO.FBParams(**kw), 2, 1, 0.48, O.LSParams(4.0))[:2]
    U, V = G.genericPyramidalOpticalFlow(a, b, 2.0, FB.Farneback_PyCL(**kw), 2, 1, 0.48, LS.LiuShenOpticalFlowAlgoAdapter(4.0))
    d = max(np.abs(U - Uo).max(), np.abs(V - Vo).max())
    print("driver FB + LS: max|d| %.3g" % d)
    assert d <= 1e-4
    # batched: a stack of pairs gives the same flows as pair-by-pair calls
    A = np.stack([a, b])
    B = np.stack([b, a])
    Ub, Vb = G.genericPyramidalOpticalFlowBatch(A, B, 2.0, FB.Farneback_PyCL(**kw), 2, 1, 0.48,
                                                  LS.LiuShenOpticalFlowAlgoAdapter(4.0))
    assert np.array_equal(Ub[0], U) and np.array_equal(Vb[0], V)


def test_farneback_argument_errors(ofri, h, mods):
    FB = mods[0]
    z = np.zeros((64, 64), np.float32)
    with pytest.raises(AssertionError):
        FB.Farneback_PyCL(polyN=6).compute(z, z, z, z)
    p = FB.Farneback_PyCL().native_params()
    p.poly_n = 6
    with pytest.raises(ValueError, match="polyN"):
        h.farneback_compute(z, z, None, None, p)
    p = FB.Farneback_PyCL().native_params()
    p.size = 12
    with pytest.raises(ValueError, match="ABI"):
        h.farneback_compute(z, z, None, None, p)
    h2 = ofri.Handle(0)
    try:
        params = ofri.make_params(ofri.fb_algo(), None, filter_sigma=0.0, pyramid_levels=1, k_levels=1)
        with pytest.raises(ValueError, match="ofri_set_farneback"):
            h2.pyramidal_flow(z, z, params)
        h2.set_farneback(FB.Farneback_PyCL().native_params())
        with pytest.raises(NotImplementedError, match="row-band"):
            h2.band_plan(64, 64, params, 0, 2)
    finally:
        h2.close()


@pytest.mark.parametrize("shape", [(37, 45), (16, 20), (65, 40), (131, 34)])
def test_small_and_odd_shapes(h, mods, shape):
    """Frames smaller than the window / the polynomial support (the kernels' reflect rule wraps with a modulo guard), widths
    that are not a multiple of 4, internal pyramid levels that stop at the 32-pixel limit (FB:483-489), batch of 2."""
    FB = mods[0]
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    H, W = shape
    A = rng.uniform(0, 255, (2, H, W)).astype(np.float32)
    B = np.roll(A, 1, 2) + rng.uniform(-2, 2, (2, H, W)).astype(np.float32)
    kw = dict(windowSize=13, Niters=2, polyN=5, polySigma=1.1, pyramidalLevels=3)
    U, V = h.farneback_compute(A, B, None, None, FB.Farneback_PyCL(**kw).native_params())
    for i in range(2):
        z = np.zeros((H, W), np.float32)
        uo, vo, _ = FBO.FBParams(**kw).compute(A[i], B[i], z, z)
        assert np.array_equal(U[i], uo) and np.array_equal(V[i], vo), (shape, i)
