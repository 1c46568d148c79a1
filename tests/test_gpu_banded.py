"""Row-band domain decomposition (BASELINE configs[4]) on ONE GPU: N virtual bands driven by N host threads
over the library's local communicator must reproduce the single-band result BIT FOR BIT on every owned row -- ghost
frames, ghost-row exchanges, global-coordinate resample / warp, all-reduced maxima and residuals, and the spline up-sample
from a halo of the coarse flow (windowed column solve; all-gather for bands thinner than the halo)."""
import numpy as np
import pytest

import ofri_oracle as O

pytestmark = pytest.mark.gpu
HS_DEF = dict(warping=True, bilinear=True, final_scaling=True)


@pytest.fixture(scope="module")
def ofri():
    import opticalflow_ri_b200 as o
    return o


def same(a, b, what):
    if not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        rows = np.unique(np.nonzero(a != b)[0])
        raise AssertionError("%s not bit-exact: max|d| = %g, %d px differ, rows %s..%s" %
                             (what, d.max(), np.count_nonzero(a != b), rows[:3], rows[-3:]))


CASES = {
    "ex3": lambda o: o.make_params(o.hs_algo([45, 21], 40), o.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                                   pyramid_levels=2, **HS_DEF),
    "hs_l1": lambda o: o.make_params(o.hs_algo([21], 50), filter_sigma=3.4, pyramid_levels=1, **HS_DEF),
    "hs_l3": lambda o: o.make_params(o.hs_algo([45, 30, 21], 36), filter_sigma=3.4, pyramid_levels=3, **HS_DEF),
    "ls_main": lambda o: o.make_params(o.ls_algo(0.1), filter_sigma=3.4, pyramid_levels=2),
    "ls_stop": lambda o: o.make_params(o.hs_algo([45, 21], 20), o.ls_algo(5, 60, 2e-5), filter_sigma=3.4,
                                       filter_opt_sigma=0.48, pyramid_levels=2, **HS_DEF),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("nb,shape", [(1, (128, 96)), (2, (256, 200)), (4, (512, 203))])
def test_band_invariance_local(ofri, case, nb, shape):
    from opticalflow_ri_b200 import banded
    H, W = shape
    I0, I1 = O.synthetic_piv_pair(H, W, seed=11)
    mk = lambda: CASES[case](ofri)
    h = ofri.Handle(0)
    try:
        Uref, Vref = h.pyramidal_flow(I0, I1, mk())
    finally:
        h.close()
    U, V = banded.flow_banded_local(I0, I1, mk, nb)
    same(U, Uref, "U %s nb=%d" % (case, nb))
    same(V, Vref, "V %s nb=%d" % (case, nb))


@pytest.mark.parametrize("nb,shape", [(2, (640, 136)), (4, (1536, 72))])
def test_band_spline_halo_path(ofri, nb, shape):
    """Bands thick enough (coarse rows per band >= the 73-row spline halo) take the windowed column solve fed by a halo
    exchange instead of the all-gather of the coarse flow; the result must still equal the single-band one bit for bit,
    and the all-gather fall-back (spline_variant 0) as well."""
    from opticalflow_ri_b200 import banded
    H, W = shape
    I0, I1 = O.synthetic_piv_pair(H, W, seed=5)
    mk = lambda: CASES["ex3"](ofri)
    h = ofri.Handle(0)
    try:
        Uref, Vref = h.pyramidal_flow(I0, I1, mk())
    finally:
        h.close()
    U, V = banded.flow_banded_local(I0, I1, mk, nb)
    same(U, Uref, "U halo nb=%d" % nb)
    same(V, Vref, "V halo nb=%d" % nb)
    U, V = banded.flow_banded_local(I0, I1, mk, nb, options={"spline_variant": 0})
    same(U, Uref, "U all-gather nb=%d" % nb)
    same(V, Vref, "V all-gather nb=%d" % nb)


@pytest.mark.parametrize("nb", [1, 2, 4])
def test_band_reference_golden_2048(ofri, big2048, nb):
    """The row-band path against the REFERENCE itself (not against this library's single-GPU path): one seeded synthetic
    2048 x 2048 pair, full EX3 parameters, nb virtual bands on one GPU (N = 2, 4: ghost-row exchanges every 32 sweeps,
    all-reduced Liu-Shen maxima / residuals, halo-fed windowed spline solve)."""
    from opticalflow_ri_b200 import banded
    g = big2048
    mk = lambda: CASES_FULL(ofri)
    U, V = banded.flow_banded_local(g["im0"], g["im1"], mk, nb)
    du, dv = float(np.max(np.abs(U - g["U"]))), float(np.max(np.abs(V - g["V"])))
    Ut, Vt = O.poiseuille_truth(2048, 2048)
    de = abs(O.epe_rmse(U, V, Ut, Vt) - O.epe_rmse(g["U"], g["V"], Ut, Vt))
    print("\n2048^2, %d band(s) vs reference: max|dU| %.3g max|dV| %.3g |dEPE-RMSE| %.3g" % (nb, du, dv, de))
    assert du <= 1e-4 and dv <= 1e-4 and de <= 1e-6, (du, dv, de)


def CASES_FULL(o):
    return o.make_params(o.hs_algo([45, 21], 600), o.ls_algo(5), filter_sigma=3.4, filter_opt_sigma=0.48,
                         pyramid_levels=2, **HS_DEF)


def test_band_plan_and_errors(ofri):
    h = ofri.Handle(0)
    try:
        p = CASES["ex3"](ofri)
        b = h.band_plan(1024, 1024, p, 1, 4)
        assert (b.own0, b.own1) == (256, 512) and b.in0 < b.own0 - b.ghost and b.in1 > b.own1 + b.ghost
        assert b.exchange % h.get_option("hs_fuse") == 0
        with pytest.raises(NotImplementedError):      # 1000 is not divisible by 3 ranks x 2
            h.band_plan(1000, 64, p, 0, 3)
        with pytest.raises(ValueError):               # 8 coarse rows per rank < the exchange interval
            h.band_plan(64, 64, p, 0, 4)
    finally:
        h.close()
