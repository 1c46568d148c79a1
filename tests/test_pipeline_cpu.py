"""The threading logic of the file -> GPU -> file pipeline (opticalflow_ri_b200/pipeline.py) without a GPU: decode into
ring slots, one "compute" per slot, results written in order, slots recycled, errors propagated.  The native compute is
replaced by an injected function here; the GPU test of the real thing is tests/test_gpu_parity.py::test_sequence_pipeline."""
import threading
import time

import numpy as np
import pytest

from opticalflow_ri_b200.pipeline import SequencePipeline


def fake_read(path):
    i = int(path.split("_")[1])
    return np.full((6, 8), float(i), np.float32)


def make(batch=3, ring=2, slow=0.0, fail_at=None):
    seen = []

    def compute(s):
        if slow:
            time.sleep(slow)
        if fail_at is not None and len(seen) == fail_at:
            raise RuntimeError("gpu call failed")
        seen.append((s.n, threading.get_ident()))
        s.u[:s.n] = s.im1[:s.n] + s.im2[:s.n]
        s.v[:s.n] = s.im2[:s.n] - s.im1[:s.n]

    p = SequencePipeline(None, None, 6, 8, batch=batch, ring=ring, decode_workers=2, compute_fn=compute,
                         alloc_fn=lambda shape: np.zeros(shape, np.float32), read_fn=fake_read)
    return p, seen


def test_pipeline_orders_results_and_recycles_slots():
    p, seen = make(batch=3, ring=2, slow=0.01)
    pairs = [("f_%d" % i, "f_%d" % (i + 1), "pair%02d" % i) for i in range(10)]      # consecutive frames: shared decode
    got = {}
    st = p.run(pairs, on_result=lambda name, U, V: got.__setitem__(name, (U.copy(), V.copy())))
    assert st["pairs"] == 10 and [n for n, _ in seen] == [3, 3, 3, 1]
    assert sorted(got) == ["pair%02d" % i for i in range(10)]
    for i in range(10):
        U, V = got["pair%02d" % i]
        assert np.all(U == 2 * i + 1) and np.all(V == 1)
    assert len(p.slots) == 2                                   # 4 batches went through 2 slots


def test_pipeline_writes_mat_files(tmp_path):
    pytest.importorskip("scipy.io")
    p, _ = make(batch=4, ring=3)
    st = p.run([("f_1", "f_2"), ("f_2", "f_5")], out_dir=str(tmp_path))
    assert st["pairs"] == 2 and sorted(x.name for x in tmp_path.iterdir()) == ["f_1.mat", "f_2.mat"]


def test_pipeline_propagates_errors():
    p, _ = make(batch=2, ring=2, fail_at=1)
    with pytest.raises(RuntimeError, match="gpu call failed"):
        p.run([("f_%d" % i, "f_%d" % (i + 1)) for i in range(8)], on_result=lambda *a: None)
    p, _ = make(batch=2, ring=2)
    with pytest.raises(ValueError):                            # a frame of the wrong size
        p.read = lambda path: np.zeros((5, 5), np.float32)
        p.run([("f_1", "f_2")], on_result=lambda *a: None)
