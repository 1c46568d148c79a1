"""Worker of tests/test_multirank_cpu.py (one process per rank, torch.distributed gloo, no GPU).

Checks the HOST logic of the row-band decomposition with real message passing: every rank takes its band plan from
the C ABI (ofri_band_plan_host), holds only the rows the plan asks for, runs the ORACLE's stages on its band as if it
were a whole image (Gaussian pre-filter, 2x2 derivatives, E Horn-Schunck sweeps), refreshes its ghost rows from the
neighbours' owned rows every E sweeps over gloo send / recv -- the protocol of the CUDA driver (ofri_api.cu,
run_pyramid_banded) -- and compares its owned rows bit for bit with the whole-image oracle run.

Second part: the DISTRIBUTED spline up-sample of the coarse flow.  A rank holds only its own coarse rows, receives the
halo the windowed column solve needs (ofri_spline.cuh; the row count comes from the same header through the host
harness) from its neighbours, and up-samples its band with the chunked / windowed algorithm; every coarse row it does
not hold is NaN, so reading outside the window would poison the result.  Compared bit for bit with the whole-plane
oracle spline."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ofri_oracle as O  # noqa: E402
import opticalflow_ri_b200 as ofri  # noqa: E402


def exchange(U, V, o0, o1, E, rank, world):
    """ghost rows [o0-E, o0) / [o1, o1+E) <- the neighbours' owned rows; even ranks send first (no deadlock)."""
    def send(dst, rows):
        dist.send(torch.from_numpy(np.ascontiguousarray(np.stack([U[rows], V[rows]]))), dst)

    def recv(src, rows):
        t = torch.empty((2, rows.stop - rows.start, U.shape[1]), dtype=torch.float32)
        dist.recv(t, src)
        U[rows], V[rows] = t[0].numpy(), t[1].numpy()

    for phase in (0, 1):
        if rank % 2 == phase:
            if rank + 1 < world:
                send(rank + 1, slice(o1 - E, o1))
                recv(rank + 1, slice(o1, o1 + E))
        else:
            if rank > 0:
                recv(rank - 1, slice(o0 - E, o0))
                send(rank - 1, slice(o0, o0 + E))


def _hostcheck():
    import ctypes as C
    import subprocess
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    out = os.path.join(ROOT, "tests", "hostcheck", "_build", "libhostcheck_mp.so")
    if dist.get_rank() == 0:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", out], check=True)
    dist.barrier()
    return C.CDLL(out)


def spline_part(rank, world, G):
    """coarse plane h x w = (96 world) x 37 -> fine (192 world) x 75; rank r owns coarse rows [96 r, 96 (r+1)) and
    up-samples fine rows [192 r - G, 192 (r+1) + G) (its extended band) with a (f32) scale factor."""
    import ctypes as C
    hc = _hostcheck()
    fp = C.POINTER(C.c_float)
    per_c, per_f, w, W = 96, 192, 37, 75
    h, H = per_c * world, per_f * world
    rng = np.random.default_rng(3)
    full = (np.cumsum(rng.normal(0, 0.4, (h, w)), axis=0) + rng.normal(0, 1, (h, w))).astype(np.float32)
    need = []
    for r in range(world):                                   # every rank derives the same halo (symmetric exchange)
        e0, e1 = max(0, r * per_f - G), min(H, (r + 1) * per_f + G)
        lo, hi = C.c_int(), C.c_int()
        hc.hc_spline_rows_needed(e0, e1 - e0, h, H, 48, C.byref(lo), C.byref(hi))
        need.append((e0, e1, lo.value, hi.value))
    halo = max(max(r * per_c - n[2], n[3] - (r + 1) * per_c) for r, n in enumerate(need))
    assert 0 < halo <= per_c
    own0, own1 = rank * per_c, (rank + 1) * per_c
    plane = np.full((h, w), np.nan, np.float32)
    plane[own0:own1] = full[own0:own1]                       # the only rows this rank has by itself
    # halo rows from the neighbours (same even / odd ordering as the ghost-row exchange)
    exchange_rows(plane, own0, own1, halo, rank, world)
    e0, e1, lo, hi = need[rank]
    assert max(0, own0 - halo) <= lo and hi <= min(h, own1 + halo)
    out = np.empty((e1 - e0, W), np.float32)
    mul = np.float32(np.float32(W) / np.float32(w))
    hc.hc_spline_win(plane.ctypes.data_as(fp), h, w, H, W, C.c_float(mul), e0, e1 - e0, 16, 9, 48, out.ctypes.data_as(fp))
    ref = (O.spline_upsample(full, H, W) * mul).astype(np.float32)[e0:e1]
    return int(np.count_nonzero(out != ref)) + int(np.isnan(out).sum())


def exchange_rows(P, o0, o1, n, rank, world):
    def send(dst, rows):
        dist.send(torch.from_numpy(np.ascontiguousarray(P[rows])), dst)

    def recv(src, rows):
        t = torch.empty((rows.stop - rows.start, P.shape[1]), dtype=torch.float32)
        dist.recv(t, src)
        P[rows] = t.numpy()

    for phase in (0, 1):
        if rank % 2 == phase:
            if rank + 1 < world:
                send(rank + 1, slice(o1 - n, o1))
                recv(rank + 1, slice(o1, o1 + n))
        else:
            if rank > 0:
                recv(rank - 1, slice(o0 - n, o0))
                send(rank - 1, slice(o0, o0 + n))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, NITER, ALPHA, T = 64 * world, 40, 44, 21.0, 4
    p = ofri.make_params(ofri.hs_algo([ALPHA], NITER), filter_sigma=3.4, pyramid_levels=1, warping=True, bilinear=True,
                         final_scaling=True)
    band = ofri.band_plan_host(H, W, p, rank, world, hs_fuse=T)
    plans = [None] * world
    dist.all_gather_object(plans, (band.own0, band.own1, band.in0, band.in1, band.ghost, band.exchange))
    # the bands tile the frame, every rank's input rows cover its owned rows + ghost frame, all ranks agree on E and G
    assert [q[0] for q in plans] == [r * H // world for r in range(world)] and plans[-1][1] == H
    assert all(plans[i][1] == plans[i + 1][0] for i in range(world - 1))
    assert len({q[4:] for q in plans}) == 1 and band.exchange % T == 0 and band.ghost >= band.exchange + 2
    assert band.in0 == max(0, band.own0 - band.ghost) and band.in1 == min(H, band.own1 + band.ghost)
    E = band.exchange
    I0, I1 = O.synthetic_piv_pair(H, W, seed=5)                      # same frames on every rank (seeded)
    a, b = I0[band.in0:band.in1], I1[band.in0:band.in1]                  # ... but a rank only LOOKS at its rows
    f1, f2 = O.gaussian_filter_px(a, 3.4, 3), O.gaussian_filter_px(b, 3.4, 3)
    fx, fy, ft = O.hs_derivatives(f1, f2)
    U, V = np.zeros_like(a), np.zeros_like(a)
    o0, o1 = band.own0 - band.in0, band.own1 - band.in0
    done = 0
    while done < NITER:
        n = min(T, NITER - done)
        U, V = O.hs_iterate(U, V, fx, fy, ft, ALPHA, n)
        U, V = np.ascontiguousarray(U), np.ascontiguousarray(V)
        done += n
        if done % E == 0 or done == NITER:
            exchange(U, V, o0, o1, E, rank, world)
    # whole-image oracle on the same frames
    F1, F2 = O.gaussian_filter_px(I0, 3.4, 3), O.gaussian_filter_px(I1, 3.4, 3)
    Ur, Vr, _ = O.hs_compute(F1, F2, ALPHA, NITER, np.zeros_like(I0), np.zeros_like(I0))
    bad = int(np.count_nonzero(U[o0:o1] != Ur[band.own0:band.own1]) + np.count_nonzero(V[o0:o1] != Vr[band.own0:band.own1]))
    # after the final exchange the ghost rows next to the owned rows are exact too
    g0, g1 = max(o0 - E, 0), min(o1 + E, U.shape[0])
    bad += int(np.count_nonzero(U[g0:g1] != Ur[band.in0 + g0:band.in0 + g1]))
    bad += spline_part(rank, world, band.ghost)
    t = torch.tensor([bad])
    dist.all_reduce(t)
    if rank == 0:
        print("MP_BAND_RESULT mismatching_px=%d ranks=%d exchange=%d ghost=%d" % (int(t.item()), world, E, band.ghost), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 3)


if __name__ == "__main__":
    main()
