"""Placeholder so that benchmark_of_methods.py's top-level imports resolve: the Farneback OpenCL adapter is outside
this package's scope.  Constructing it raises, which BOM's try/except turns into a skipped row."""


class Farneback_PyCL(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("Farneback (OpenCL) is not part of the B200 HS / Liu-Shen path")
