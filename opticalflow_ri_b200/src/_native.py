"""Shared plumbing of the drop-in modules: locate the package that sits one directory up and hand out the GPU handle."""
import os
import sys

_PKG_PARENT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_PARENT not in sys.path:
    sys.path.insert(0, _PKG_PARENT)

import opticalflow_ri_b200 as ofri  # noqa: E402

VERBOSE = bool(int(os.environ.get("OFRI_VERBOSE", "0")))   # the reference prints progress; the drop-in is quiet


def handle():
    return ofri.default_handle(int(os.environ.get("OFRI_DEVICE", "0")))


def log(*a):
    if VERBOSE:
        print(*a)
