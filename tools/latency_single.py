#!/usr/bin/env python
"""Steady-state latency of ONE pair through the drop-in call (numpy in, numpy out), EX3 parameters."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "opticalflow_ri_b200", "src"))
sys.path.insert(0, ROOT)
from GenericPyramidalOpticalFlow import genericPyramidalOpticalFlow  # noqa: E402
from HornSchunck import HSOpticalFlowAlgoAdapter  # noqa: E402
from PhysicsBasedOpticalFlowLiuShen import LiuShenOpticalFlowAlgoAdapter  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

import opticalflow_ri_b200 as ofri  # noqa: E402
_h = ofri.default_handle(0)
for k in ("hs_fuse", "hs_variant", "ls_fuse", "ls_variant"):
    if os.environ.get(k.upper()):
        _h.set_option(k, int(os.environ[k.upper()]))
print({k: _h.get_option(k) for k in ("hs_fuse", "hs_variant", "ls_fuse", "ls_variant")})
for n in (512, 1024, 2048):
    a, b = synthetic_piv_pair(n, n, 0)
    ts = []
    for i in range(6):
        t = time.perf_counter()
        genericPyramidalOpticalFlow(a, b, 3.4, HSOpticalFlowAlgoAdapter([21, 45], 600), 2, 1, 0.48, LiuShenOpticalFlowAlgoAdapter(5))
        ts.append(time.perf_counter() - t)
    print("single pair %4d x %-4d: first %.1f ms, steady %.2f ms (min of 5)" % (n, n, ts[0] * 1e3, min(ts[1:]) * 1e3), flush=True)
