#!/usr/bin/env python
"""Small end-to-end + banded run for compute-sanitizer (memcheck): every kernel family once, border tiles included."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200 import banded  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402

h = ofri.Handle(0)
mk = lambda: ofri.make_params(ofri.hs_algo([45.0, 21.0], 12), ofri.ls_algo(5.0, 6), filter_sigma=3.4, filter_opt_sigma=0.48,
                              pyramid_levels=2, warping=True, bilinear=True, final_scaling=True)
for shape in ((150, 203), (256, 384)):
    a, b = synthetic_piv_pair(*shape, seed=1)
    U, V = h.pyramidal_flow(np.stack([a, a, b]), np.stack([b, b, a]), mk())
    assert np.isfinite(U).all() and np.isfinite(V).all()
    for prec in (0, 2):
        h.set_option("hs_precise", prec)
        h.pyramidal_flow(a, b, mk())
    h.set_option("hs_precise", 1)
a, b = synthetic_piv_pair(256, 200, seed=2)
U1, V1 = h.pyramidal_flow(a, b, mk())
U2, V2 = banded.flow_banded_local(a, b, mk, 2)
assert np.array_equal(U1, U2) and np.array_equal(V1, V2)
print("sanitize_small: ok")
