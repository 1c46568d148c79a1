#!/usr/bin/env python
"""Where does a tile's time go inside the persistent TMA kernels?  Builds / loads libofri_phase.so (the
-DOFRI_PHASE_TIMING variant: thread 0 of every CTA accumulates clock64() deltas per phase), runs one pipeline pass
(P pairs of SIZE^2, EX3 parameters) per setting and prints, per kernel family, the share of CTA time per phase.

  OFRI_LIB=opticalflow_ri_b200/libofri_phase.so python tools/phase_timing.py          (after build.py --phase)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OFRI_LIB", os.path.join(ROOT, "opticalflow_ri_b200", "libofri_phase.so"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

P = int(os.environ.get("SWEEP_PAIRS", "64"))
H = W = int(os.environ.get("SWEEP_SIZE", "1024"))
HS_NAMES = ["tma_wait", "ghosts", "stage_to_regs", "sweeps", "stores", "decode", "cluster_barrier", "partner_wait(in sweeps)"]
LS_NAMES = ["tma_wait", "-", "stage_to_regs", "sweeps+stores", "final_barrier", "bookkeeping", "-", "-"]
h = ofri.Handle(0)
base = [synthetic_piv_pair(H, W, s) for s in range(4)]
a = torch.from_numpy(np.stack([base[i % 4][0] for i in range(P)])).cuda()
b = torch.from_numpy(np.stack([base[i % 4][1] for i in range(P)])).cuda()
u = torch.empty_like(a)
v = torch.empty_like(a)
params = ofri.make_params(ofri.hs_algo([45.0, 21.0], 600), ofri.ls_algo(5.0, 60), filter_sigma=3.4, filter_opt_sigma=0.48,
                          pyramid_levels=2, warping=True, bilinear=True, final_scaling=True)


def run(tag, **opts):
    for k, val in opts.items():
        h.set_option(k, val)
    h.set_option("timing", 0)
    h.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), P, H, W, params, u.data_ptr(), v.data_ptr(), None, device=True)
    h.synchronize()
    h.phase_cycles(0), h.phase_cycles(1)          # reset
    h.set_option("timing", 1)
    h.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), P, H, W, params, u.data_ptr(), v.data_ptr(), None, device=True)
    h.synchronize()
    st = h.stage_timings()
    out = {"tag": tag, "opts": opts, "stage_ms": {k: round(x, 2) for k, x in st.items() if "iterate" in k}}
    for fam, names in ((0, HS_NAMES), (1, LS_NAMES)):
        c = h.phase_cycles(fam)
        tot = float(sum(c)) or 1.0
        out["hs" if fam == 0 else "ls"] = {n: round(x / tot, 4) for n, x in zip(names, c) if n != "-"}
        out[("hs" if fam == 0 else "ls") + "_Mcycles_per_cta"] = round(tot / 148 / 1e6, 3)
    print(json.dumps(out), flush=True)


for spec in os.environ.get("PHASE_RUNS", "default;hs_T8:hs_fuse=8;hs_precise_everywhere:hs_precise=2").split(";"):
    tag, _, kv = spec.partition(":")
    run(tag, **{k: int(v) for k, v in (x.split("=") for x in kv.split(",") if x)})
