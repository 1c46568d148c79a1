"""Placeholder so that benchmark_of_methods.py's top-level imports resolve: the dense Lucas-Kanade OpenCL adapter is
outside this package's scope.  Constructing it raises, which BOM's try/except turns into a skipped row."""


class denseLucasKanade_PyCl(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("dense Lucas-Kanade (OpenCL) is not part of the B200 HS / Liu-Shen path")
