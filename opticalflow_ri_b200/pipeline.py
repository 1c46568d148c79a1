"""File -> GPU -> file pipeline for PIV sequences: what surrounds the hot path in the reference's scripts
(examples/PyHSchunck_Fs3_4.py:129-141: imread both frames, compute, save_flow) restructured so that none of it sits on
the GPU's critical path once a pair takes milliseconds instead of seconds:

   decode threads           compute thread                     writer thread
   TIFF / PNG / .npy  --->  ring slot (page-locked memory) ---> one native call per slot ---> .mat per pair
   (Pillow, GIL released)   ofri_pyramidal_flow: H2D || kernels || D2H inside the call        (scipy.io.savemat)

A ring of `ring` slots, each holding `batch` pairs (inputs and outputs) in page-locked host memory, decouples the three
stages: while the GPU works on slot k the decoders fill slot k+1 and the writer drains slot k-1.  All arithmetic is in
libofri.so; nothing here computes flow.  `compute_fn` / `alloc_fn` are injectable so that the threading logic is
testable without a GPU."""
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import io as _io


class _Slot(object):
    def __init__(self, alloc, batch, H, W):
        self.im1 = alloc((batch, H, W))
        self.im2 = alloc((batch, H, W))
        self.u = alloc((batch, H, W))
        self.v = alloc((batch, H, W))
        self.n = 0
        self.names = []


class SequencePipeline(object):
    """pipeline = SequencePipeline(handle, params_factory, H, W); stats = pipeline.run(pairs, out_dir)

    pairs: list of (path_frame_a, path_frame_b[, name]); consecutive pairs that share a frame decode it once.
    params_factory(): a fresh ofri_params per native call (the HS alpha list is per call).
    on_result(name, U, V): called by the writer thread instead of writing .mat files when given."""

    def __init__(self, handle, params_factory, H, W, batch=32, ring=3, decode_workers=4, compute_fn=None, alloc_fn=None,
                 read_fn=None):
        self.h, self.mk, self.H, self.W = handle, params_factory, int(H), int(W)
        self.batch, self.ring = int(batch), max(2, int(ring))
        self.decode_workers = max(1, int(decode_workers))
        self.read = read_fn or _io.read_frame
        alloc = alloc_fn or (lambda shape: handle.pinned_empty(shape, np.float32))
        self.slots = [_Slot(alloc, self.batch, self.H, self.W) for _ in range(self.ring)]
        self.compute = compute_fn or self._native_compute

    def _native_compute(self, s):
        self.h.pyramidal_flow_ptr(s.im1.ctypes.data, s.im2.ctypes.data, s.n, self.H, self.W, self.mk(), s.u.ctypes.data,
                                  s.v.ctypes.data, None, device=False)

    def run(self, pairs, out_dir=None, on_result=None):
        pairs = [(p[0], p[1], p[2] if len(p) > 2 else os.path.splitext(os.path.basename(p[0]))[0]) for p in pairs]
        free_q, full_q, done_q = queue.Queue(), queue.Queue(), queue.Queue()
        for s in self.slots:
            free_q.put(s)
        errors = []
        t_dec = [0.0]
        t_gpu = [0.0]
        t_wr = [0.0]
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)

        def check_shape(a, path):
            if a.shape != (self.H, self.W):
                raise ValueError("%s is %r, the pipeline was built for %r" % (path, a.shape, (self.H, self.W)))
            return a

        def decoder():
            try:
                with ThreadPoolExecutor(self.decode_workers) as pool:
                    for b0 in range(0, len(pairs), self.batch):
                        chunk = pairs[b0:b0 + self.batch]
                        s = free_q.get()
                        if s is None:
                            return
                        t0 = time.perf_counter()
                        uniq = {}
                        for pa, pb, _ in chunk:
                            uniq.setdefault(pa, None)
                            uniq.setdefault(pb, None)
                        for path, arr in zip(uniq, pool.map(self.read, list(uniq))):
                            uniq[path] = check_shape(arr, path)
                        for i, (pa, pb, name) in enumerate(chunk):
                            s.im1[i] = uniq[pa]
                            s.im2[i] = uniq[pb]
                        s.n = len(chunk)
                        s.names = [c[2] for c in chunk]
                        t_dec[0] += time.perf_counter() - t0
                        full_q.put(s)
            except BaseException as e:      # noqa: BLE001
                errors.append(e)
            finally:
                full_q.put(None)

        def writer():
            try:
                while True:
                    s = done_q.get()
                    if s is None:
                        return
                    t0 = time.perf_counter()
                    for i, name in enumerate(s.names):
                        if on_result is not None:
                            on_result(name, s.u[i], s.v[i])
                        elif out_dir:
                            _io.save_flow(s.u[i], s.v[i], os.path.join(out_dir, name + ".mat"))
                    t_wr[0] += time.perf_counter() - t0
                    free_q.put(s)
            except BaseException as e:      # noqa: BLE001
                errors.append(e)
                free_q.put(None)            # unblock the decoder

        td = threading.Thread(target=decoder, daemon=True)
        tw = threading.Thread(target=writer, daemon=True)
        t_all = time.perf_counter()
        td.start()
        tw.start()
        npairs = 0
        try:
            while True:
                s = full_q.get()
                if s is None or errors:
                    break
                t0 = time.perf_counter()
                self.compute(s)             # one native call: H2D, all kernels and D2H of the slot, overlapped inside
                t_gpu[0] += time.perf_counter() - t0
                npairs += s.n
                done_q.put(s)
        except BaseException as e:          # noqa: BLE001
            errors.append(e)
        finally:
            done_q.put(None)
            free_q.put(None)
            tw.join()
            td.join(timeout=30)
        wall = time.perf_counter() - t_all
        if errors:
            raise errors[0]
        return {"pairs": npairs, "wall_s": wall, "pairs_per_s": npairs / wall if wall > 0 else 0.0,
                "decode_busy_s": t_dec[0], "gpu_call_busy_s": t_gpu[0], "writer_busy_s": t_wr[0],
                "batch": self.batch, "ring": self.ring, "decode_workers": self.decode_workers}
