"""Row-band domain decomposition of ONE very large frame pair (BASELINE.json configs[4]): host-side helpers.

`flow_banded_local` drives N bands from N host threads of this process over the library's "local" communicator (any
number of bands per GPU) -- the band-invariance test harness and the single-process multi-GPU mode.
`flow_banded_rank` is the per-rank call of the one-process-per-GPU (torchrun / NCCL) mode: every rank passes the rows
of the frames `band_plan` asks for and gets its owned rows of the flow back.  All arithmetic is in libofri.so."""
import threading

import numpy as np

from . import api


def flow_banded_rank(handle, im1_rows, im2_rows, H, W, params, band):
    """im*_rows: torch CUDA float32 tensors holding rows [band.in0, band.in1) of the frames.  Returns (U, V) torch
    tensors with rows [band.own0, band.own1)."""
    import torch
    assert im1_rows.is_cuda and im1_rows.dtype == torch.float32 and im1_rows.is_contiguous()
    assert tuple(im1_rows.shape) == (band.in1 - band.in0, W) == tuple(im2_rows.shape)
    u = torch.empty((band.own1 - band.own0, W), dtype=torch.float32, device=im1_rows.device)
    v = torch.empty_like(u)
    handle.pyramidal_flow_banded_ptr(im1_rows.data_ptr(), im2_rows.data_ptr(), H, W, params, u.data_ptr(), v.data_ptr())
    return u, v


def flow_banded_local(im1, im2, params_factory, nbands, devices=None, options=None):
    """Whole frames as numpy (H, W) arrays; `nbands` bands, band r on devices[r % len(devices)] (default: all on GPU 0).
    params_factory() must return a fresh ofri_params per band.  Returns (U, V) numpy arrays of the whole frame."""
    import torch
    im1 = np.ascontiguousarray(im1, dtype=np.float32)
    im2 = np.ascontiguousarray(im2, dtype=np.float32)
    H, W = im1.shape
    devices = list(devices) if devices else [0]
    group = api.LocalGroup(nbands)
    U = np.empty((H, W), np.float32)
    V = np.empty((H, W), np.float32)
    errors = [None] * nbands

    def worker(r):
        try:
            dev = devices[r % len(devices)]
            torch.cuda.set_device(dev)
            h = api.Handle(dev)
            for k, val in (options or {}).items():
                h.set_option(k, val)
            h.comm_init_local(group.ptr, r)
            p = params_factory()
            band = h.band_plan(H, W, p, r, nbands)
            a = torch.from_numpy(im1[band.in0:band.in1]).to("cuda:%d" % dev)
            b = torch.from_numpy(im2[band.in0:band.in1]).to("cuda:%d" % dev)
            torch.cuda.synchronize(dev)
            u, v = flow_banded_rank(h, a, b, H, W, p, band)
            U[band.own0:band.own1] = u.cpu().numpy()
            V[band.own0:band.own1] = v.cpu().numpy()
            h.comm_destroy()
            h.close()
        except BaseException as e:      # noqa: BLE001 -- reported to the caller below
            errors[r] = e
            group.abort()               # the other bands may be blocked in a collective waiting for this one

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(nbands)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    group.close()
    first = [e for e in errors if e is not None and "aborted" not in str(e)] or [e for e in errors if e is not None]
    if first:
        raise first[0]                  # the original failure, not the peers' "group aborted"
    return U, V
