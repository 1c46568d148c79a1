#!/usr/bin/env python
"""The method table of the reference's benchmark_of_methods.py (BOM:143-302) against the B200 drop-in modules: the same
ten configurations -- four Horn-Schunck rows (BOM:143-194), three dense Lucas-Kanade rows (BOM:196-248), three Farneback
rows (BOM:250-302) -- built with the same adapter constructors and run through GenericPyramidalOpticalFlowWrapper exactly
as the reference does (including its quirk that a `use_liu_shen` row runs Liu-Shen ALONE as the main adapter).  Frames
are read with Pillow instead of skimage, flows are saved in the reference's .mat layout, plots are not made.

  python examples/run_benchmark_of_methods.py [frame0.tif frame1.tif] [--out benchmark_results]

Without frame arguments the bundled 512 x 512 Poiseuille pair (tests/golden/bundled_pair.npz) is used."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opticalflow_ri_b200", "src"))
sys.path.insert(0, ROOT)

from GenericPyramidalOpticalFlowWrapper import GenericPyramidalOpticalFlowWrapper  # noqa: E402
from HornSchunck import HSOpticalFlowAlgoAdapter  # noqa: E402
from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter  # noqa: E402
from Farneback_PyCL import Farneback_PyCL  # noqa: E402
from denseLucasKanade_PyCL import denseLucasKanade_PyCl  # noqa: E402

from opticalflow_ri_b200.io import read_frame, save_flow  # noqa: E402

ROWS = [   # (name, family, filter_sigma, pyr_levels, use_liu_shen)
    ("HS_Fs0_0", "hs", 0.0, 1, False), ("HS_Fs3_4", "hs", 3.4, 1, False), ("HS_Fs3_4_PyrLvls2", "hs", 3.4, 2, False),
    ("LiuSE_HS_Fs3_4_PyrLvls2", "hs", 3.4, 2, True),
    ("LK_Fs2_0", "lk", 2.0, 1, False), ("LK_Fs2_0_PyrLvls2", "lk", 2.0, 2, False), ("LiuSE_LK_Fs2_0_PyrLvls2", "lk", 2.0, 2, True),
    ("FB_Fs0_0", "fb", 0.0, 1, False), ("FB_Fs0_0_PyrLvls2", "fb", 0.0, 2, False), ("LiuSE_FB_Fs0_0_PyrLvls2", "fb", 0.0, 2, True),
]


def make_adapter(family, pyr_levels, use_liu_shen):
    if use_liu_shen:
        return LiuShenOpticalFlowAlgoAdapter(0.1)
    if family == "hs":
        return HSOpticalFlowAlgoAdapter(alphas=[1.0] * pyr_levels, Niter=100, provideGenericPyramidalDefaults=True)
    if family == "lk":
        return denseLucasKanade_PyCl(halfWindow=13, Niter=5)
    return Farneback_PyCL(windowSize=33, Niters=5, polyN=7, polySigma=1.5)


def run_benchmark(img1, img2, output_dir=None, rows=ROWS):
    results = {}
    for name, family, sigma, levels, ls in rows:
        flow = GenericPyramidalOpticalFlowWrapper(make_adapter(family, levels, ls), filter_sigma=sigma, pyr_levels=levels)
        t = time.time()
        U, V = flow.calculateFlow(img1, img2)
        dt = time.time() - t
        results[name] = {"U": U, "V": V, "time": dt}
        print("%-26s %7.3f s   U %.2f .. %.2f   V %.2f .. %.2f" % (name, dt, U.min(), U.max(), V.min(), V.max()))
        if output_dir:
            save_flow(U, V, os.path.join(output_dir, name + ".mat"))
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("frames", nargs="*")
    ap.add_argument("--out", default="benchmark_results")
    args = ap.parse_args()
    if len(args.frames) == 2:
        img1, img2 = read_frame(args.frames[0]), read_frame(args.frames[1])
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_pair.npz"))
        img1, img2 = g["im0"].astype(np.float32), g["im1"].astype(np.float32)
    if img1.max() > 255:                      # BOM:127-130
        img1 = (img1 / 65535.0 * 255.0).astype(np.float32)
        img2 = (img2 / 65535.0 * 255.0).astype(np.float32)
    os.makedirs(args.out, exist_ok=True)
    run_benchmark(img1, img2, args.out)


if __name__ == "__main__":
    main()
