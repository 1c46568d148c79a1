#!/usr/bin/env python
"""GPU tuning sweep: per-stage device time of one full pipeline pass (P pairs of 1024x1024, EX3 parameters) for every
(hs_fuse T, tile variant) and ls_fuse, using the library's own CUDA-event stage timers.  Prints one JSON line per setting."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

P = int(os.environ.get("SWEEP_PAIRS", "64"))
H = W = int(os.environ.get("SWEEP_SIZE", "1024"))
h = ofri.Handle(0)
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
h.set_stream(_stream.cuda_stream)
base = [synthetic_piv_pair(H, W, s) for s in range(4)]
a = torch.from_numpy(np.stack([base[i % 4][0] for i in range(P)])).cuda()
b = torch.from_numpy(np.stack([base[i % 4][1] for i in range(P)])).cuda()
u = torch.empty_like(a)
v = torch.empty_like(a)
params = ofri.make_params(ofri.hs_algo([45.0, 21.0], 600), ofri.ls_algo(5.0, 60), filter_sigma=3.4, filter_opt_sigma=0.48,
                          pyramid_levels=2, warping=True, bilinear=True, final_scaling=True)
px = 1.25 * H * W * P


def run(tag):
    h.set_option("timing", 0)
    h.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), P, H, W, params, u.data_ptr(), v.data_ptr(), None, device=True)
    torch.cuda.synchronize()
    h.set_option("timing", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), P, H, W, params, u.data_ptr(), v.data_ptr(), None, device=True)
    e1.record()
    torch.cuda.synchronize()
    h.synchronize()
    st = h.stage_timings()
    tot = e0.elapsed_time(e1)
    hs = st.get("hs_iterate", 0.0) + st.get("hs_iterate_coarse", 0.0) + st.get("hs_iterate_precise", 0.0)
    ls = st.get("ls_iterate", 0.0)
    out = dict(tag, total_ms=round(tot, 2), pairs_per_s=round(P / (tot / 1e3), 1),
               hs_ms=round(hs, 2), hs_gpix_it_s=round(px * 600 / (hs / 1e3) / 1e9, 1) if hs else None,
               hs_GBs_per_T1=round(28 * px * 600 / (hs / 1e3) / 1e9, 1) if hs else None,
               ls_ms=round(ls, 2), ls_gpix_it_s=round(px * 60 / (ls / 1e3) / 1e9, 1) if ls else None,
               other_ms=round(sum(x for k, x in st.items() if k not in ("hs_iterate", "hs_iterate_coarse", "hs_iterate_precise", "ls_iterate")), 2),
               stages={k: round(x, 2) for k, x in st.items()})
    print(json.dumps(out), flush=True)


h.set_option("hs_precise", int(os.environ.get("HS_PRECISE", "1")))
for _k in ("hs_fuse_fast", "hs_fuse_precise", "auto_fuse"):
    if os.environ.get(_k.upper()):
        h.set_option(_k, int(os.environ[_k.upper()]))
which = os.environ.get("SWEEP", "hs,ls")
if "one" in which:
    h.set_option("hs_fuse", int(os.environ.get("HS_FUSE", "4")))
    h.set_option("hs_variant", int(os.environ.get("HS_VARIANT", "0")))
    h.set_option("ls_fuse", int(os.environ.get("LS_FUSE", "2")))
    h.set_option("ls_variant", int(os.environ.get("LS_VARIANT", "0")))
    run({"hs_fuse": h.get_option("hs_fuse"), "hs_variant": h.get_option("hs_variant"), "ls_fuse": h.get_option("ls_fuse")})
if "hs" in which:
    h.set_option("ls_fuse", 2)
    for T in [int(x) for x in os.environ.get("HS_TS", "0,1,2,3,4,6").split(",")]:
        for variant in ([int(x) for x in os.environ.get("HS_VARIANTS", "0,1,2,3,4,5,6,7").split(",")] if T > 0 else [0]):
            h.set_option("hs_fuse", T)
            h.set_option("hs_variant", variant)
            run({"hs_fuse": T, "hs_variant": variant, "ls_fuse": 2})
if "ls" in which:
    h.set_option("hs_fuse", 4)
    h.set_option("hs_variant", int(os.environ.get("HS_VARIANT", "0")))
    for T in [int(x) for x in os.environ.get("LS_TS", "0,1,2,3,4").split(",")]:
        for lv in ([int(x) for x in os.environ.get("LS_VARIANTS", "0,1,2,3,4,5,8").split(",")] if T > 0 else [0]):
            h.set_option("ls_fuse", T)
            h.set_option("ls_variant", lv)
            run({"hs_fuse": 4, "ls_fuse": T, "ls_variant": lv})
