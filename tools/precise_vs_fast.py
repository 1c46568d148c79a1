#!/usr/bin/env python
"""How far does the fast Horn-Schunck arithmetic (hs_precise = 0) move the final flow away from the reference arithmetic
(hs_precise = 2, bit-exact on every level), as a function of the regularisation weight alpha?  Bundled Poiseuille pair
(512 x 512) and a synthetic 1024 x 1024 pair, 2 pyramid levels, FILTER 3.4, 600 (EX) / 100 (BOM) sweeps, with and
without the Liu-Shen refinement.  Prints one JSON line per case: max |dU|, max |dV| and the EPE-RMSE difference against
the analytic Poiseuille profile.  Basis of the `hs_precise = 1` rule."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_pair.npz"))
pairs = {"bundled512": (g["im0"].astype(np.float32), g["im1"].astype(np.float32)),
         "synthetic1024": synthetic_piv_pair(1024, 1024, 3)}
h = ofri.Handle(0)


def truth(H, W):
    y = np.arange(H, dtype=np.float64)[:, None]
    u = -4.0 * (1.0 - ((y - (H - 1) / 2.0) / (H / 2.0)) ** 2)
    return np.broadcast_to(u, (H, W)), np.zeros((H, W))


for name, (i0, i1) in pairs.items():
    H, W = i0.shape
    tu, tv = truth(H, W)
    for alpha in (0.5, 1, 2, 3, 5, 8, 10, 15, 21, 45):
        for niter, ls in ((600, True), (600, False), (100, False)):
            mk = lambda: ofri.make_params(ofri.hs_algo([float(alpha), float(alpha)], niter),
                                          ofri.ls_algo(5.0, 60) if ls else None, filter_sigma=3.4,
                                          filter_opt_sigma=0.48 if ls else None, pyramid_levels=2, warping=True,
                                          bilinear=True, final_scaling=True)
            out = {}
            for mode in (2, 1, 0):
                h.set_option("hs_precise", mode)
                out[mode] = h.pyramidal_flow(i0, i1, mk())
            rm = {m: float(np.sqrt(np.mean((out[m][0] - tu) ** 2 + (out[m][1] - tv) ** 2))) for m in out}
            print(json.dumps({"pair": name, "alpha": alpha, "niter": niter, "liu_shen": ls,
                              "max_dU_fast": float(np.abs(out[0][0] - out[2][0]).max()),
                              "max_dV_fast": float(np.abs(out[0][1] - out[2][1]).max()),
                              "max_dU_coarse_precise": float(np.abs(out[1][0] - out[2][0]).max()),
                              "max_dV_coarse_precise": float(np.abs(out[1][1] - out[2][1]).max()),
                              "d_epe_rmse_fast": abs(rm[0] - rm[2]), "d_epe_rmse_coarse_precise": abs(rm[1] - rm[2])}),
                  flush=True)
h.set_option("hs_precise", 1)
