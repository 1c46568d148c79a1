#!/usr/bin/env python
"""Stage times of the Farneback and dense Lucas-Kanade adapters at the reference examples' parameters,
as main adapter of a one-level driver call on a batch of synthetic PIV pairs (library stage timers, CUDA events)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opticalflow_ri_b200", "src"))
sys.path.insert(0, ROOT)
from Farneback_PyCL import Farneback_PyCL  # noqa: E402
from denseLucasKanade_PyCL import denseLucasKanade_PyCl  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402
import opticalflow_ri_b200 as ofri  # noqa: E402

h = ofri.Handle(0)
h.set_option("timing", 1)
REPS = int(os.environ.get("ADAPTER_REPS", "4"))
SIZES = [(int(a), int(b)) for a, b in (x.split("x") for x in os.environ.get("ADAPTER_SIZES", "512x16,1024x8").split(","))]
for n, batch in SIZES:
    pairs = [synthetic_piv_pair(n, n, s) for s in range(batch)]
    A = np.stack([p[0] for p in pairs])
    B = np.stack([p[1] for p in pairs])
    h.set_farneback(Farneback_PyCL().native_params())
    h.set_lk(denseLucasKanade_PyCl(Niter=5, halfWindow=13).native_params())
    for name, algo, sigma in (("farneback", ofri.fb_algo(), 0.0), ("lucas_kanade", ofri.lk_algo(), 2.0)):
        p = ofri.make_params(algo, None, filter_sigma=sigma, pyramid_levels=1, k_levels=1, warping=False, final_scaling=False)
        best = None
        for rep in range(REPS):
            h.pyramidal_flow(A, B, p)
            ms = h.stage_timings().get(name, 0.0)
            best = ms if best is None else min(best, ms)
        print(json.dumps({"adapter": name, "size": n, "pairs": batch, "ms_per_pair": round(best / batch, 3),
                          "Mpx_per_s": round(n * n * batch / best / 1e3, 1)}), flush=True)
