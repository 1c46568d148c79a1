#!/usr/bin/env python
"""One very large frame pair (BASELINE configs[4] shape: 16384 x 16384) through the whole path on ONE GPU, device
resident, with the per-stage CUDA-event breakdown.  The pair is a 1024 x 1024 seeded synthetic PIV pair tiled to the
requested size (content does not matter for timing).  Prints one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

N = int(os.environ.get("BIG_SIZE", "16384"))
REPS = int(os.environ.get("BIG_REPS", "2"))
h = ofri.Handle(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
h.set_stream(stream.cuda_stream)
for k in ("hs_fuse", "hs_variant", "hs_precise", "ls_fuse", "ls_variant"):
    if os.environ.get(k.upper()):
        h.set_option(k, int(os.environ[k.upper()]))
t0, t1 = synthetic_piv_pair(1024, 1024, 0)
r = N // 1024
a = torch.from_numpy(t0).cuda().repeat(r, r).contiguous().view(1, N, N)
b = torch.from_numpy(t1).cuda().repeat(r, r).contiguous().view(1, N, N)
u = torch.empty_like(a)
v = torch.empty_like(a)
params = ofri.make_params(ofri.hs_algo([45.0, 21.0], 600), ofri.ls_algo(5.0, 60), filter_sigma=3.4, filter_opt_sigma=0.48,
                          pyramid_levels=2, warping=True, bilinear=True, final_scaling=True)
best = None
for rep in range(REPS + 1):
    h.set_option("timing", 1 if rep == REPS else 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h.pyramidal_flow_ptr(a.data_ptr(), b.data_ptr(), 1, N, N, params, u.data_ptr(), v.data_ptr(), None, device=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if rep > 0:
        best = ms if best is None else min(best, ms)
h.synchronize()
st = h.stage_timings()
px_it = 1.25 * N * N * 660
print(json.dumps({"size": N, "ms": round(best, 1), "pairs_per_s": round(1e3 / best, 4),
                  "gpix_iter_per_s": round(px_it / (best / 1e3) / 1e9, 1),
                  "mem_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                  "free_total_GiB": [round(x / 2 ** 30, 1) for x in torch.cuda.mem_get_info()],
                  "finite": bool(torch.isfinite(u).all().item()), "u_mean": float(u.mean().item()),
                  "stages": {k: round(x, 1) for k, x in st.items()}}))
