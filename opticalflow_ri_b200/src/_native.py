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


_aux = None


def aux_handle():
    """A second handle on the same device for adapters whose compute() can be called back from inside a native call on
    the default handle (a handle is not re-entrant)."""
    global _aux
    if _aux is None:
        _aux = ofri.Handle(int(os.environ.get("OFRI_DEVICE", "0")))
    return _aux


def log(*a):
    if VERBOSE:
        print(*a)
