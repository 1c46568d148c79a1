#!/usr/bin/env python
"""The example configurations of the reference against the B200 drop-in modules: the same adapter constructors and the
same positional genericPyramidalOpticalFlow call, frames read with Pillow instead of skimage, the result saved in the
reference's .mat layout.  The three BASELINE configurations (examples/PyHSchunck_Fs3_4.py, PyHSchunck_Fs3_4_PyrLvls2.py,
LiuSE_PyHSchunck_Fs3_4_PyrLvls2.py) plus the Farneback and dense Lucas-Kanade ones (examples/Farneback_Fs0_0.py,
Farneback_Fs0_0_PyrLvls2.py, LiuSE_Farneback_Fs0_0_PyrLvls2.py, denseLK_Fs2_0.py, LiuSE_denseLK_Fs2_0_PyrLvls2.py).

  python examples/run_example.py CONFIG [frame0.tif frame1.tif] [--out flow.mat]

Without frame arguments the bundled 512 x 512 Poiseuille pair (tests/golden/bundled_pair.npz) is used."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opticalflow_ri_b200", "src"))      # where the reference's scripts put '../src'
sys.path.insert(0, ROOT)

from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow  # noqa: E402
from HornSchunck import HSOpticalFlowAlgoAdapter  # noqa: E402
from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter  # noqa: E402
from Farneback_PyCL import Farneback_PyCL  # noqa: E402
from denseLucasKanade_PyCL import denseLucasKanade_PyCl  # noqa: E402

from opticalflow_ri_b200.io import read_frame, save_flow  # noqa: E402

# Horn-Schunck regularisation per pyramid level for the Ni06 / Bits08 images (the reference's table: 21 at the finest
# level, 45 on coarser ones); the adapter pops from the END, so the finest level comes first in the list
CONFIGS = {
    "hs": dict(main="hs", levels=1, alphas=[21], FILTER=3.4, liu_shen=None),
    "hs_pyr2": dict(main="hs", levels=2, alphas=[21, 45], FILTER=3.4, liu_shen=None),
    "liuse_hs_pyr2": dict(main="hs", levels=2, alphas=[21, 45], FILTER=3.4, liu_shen=5),
    "farneback": dict(main="fb", levels=1, FILTER=0, liu_shen=None),
    "farneback_pyr2": dict(main="fb", levels=2, FILTER=0, liu_shen=None),
    "liuse_farneback_pyr2": dict(main="fb", levels=2, FILTER=0, liu_shen=10),
    "denselk": dict(main="lk", levels=1, FILTER=2, liu_shen=None, kw=dict(warping=False)),
    "liuse_denselk_pyr2": dict(main="lk", levels=2, FILTER=2, liu_shen=10, kw=dict(warping=False)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=sorted(CONFIGS))
    ap.add_argument("frames", nargs="*")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if len(args.frames) == 2:
        Iold, Inew = read_frame(args.frames[0]), read_frame(args.frames[1])
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_pair.npz"))
        Iold, Inew = g["im0"].astype(np.float32), g["im1"].astype(np.float32)
    FILTER, FILTER_OPT, kLevels = cfg["FILTER"], 0.48, 1
    if cfg["main"] == "hs":
        mainAdapter = HSOpticalFlowAlgoAdapter(list(cfg["alphas"]), 600)
    elif cfg["main"] == "fb":
        mainAdapter = Farneback_PyCL(platformID=0)
    else:
        mainAdapter = denseLucasKanade_PyCl(Niter=5, halfWindow=13, platformID=0)
    lsAdapter = LiuShenOpticalFlowAlgoAdapter(cfg["liu_shen"]) if cfg["liu_shen"] else None
    t = time.time()
    [U, V] = genericPyramidalOpticalFlow(Iold, Inew, FILTER, mainAdapter, cfg["levels"], kLevels, FILTER_OPT, lsAdapter,
                                         **cfg.get("kw", {}))
    dt = time.time() - t
    print("%s: %d x %d, %.3f s, U in [%.4f, %.4f], V in [%.4f, %.4f]" % (args.config, U.shape[0], U.shape[1], dt,
                                                                        U.min(), U.max(), V.min(), V.max()))
    out = args.out or os.path.join(".", "%s.mat" % args.config)
    save_flow(U, V, out)
    print("saved", out)


if __name__ == "__main__":
    main()
