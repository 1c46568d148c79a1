#!/usr/bin/env python
"""One very large frame pair over N GPUs by row-band domain decomposition (BASELINE configs[4]): one process per GPU
under torchrun, NCCL halo exchange inside libofri.so.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/run_banded.py [--size 16384] [--check-size 2048] [--reps 2]

Each rank builds the rows of the frames it needs (a seeded 1024 x 1024 synthetic PIV pair tiled to size x size), runs
the banded path, and rank 0 prints one JSON line: time per pair (CUDA events, max over ranks; median of --reps
repetitions), pairs/s, Gpix-sweeps/s.
--check-size > 0 first verifies, at that size, that every rank's owned rows are bit-identical to the single-GPU path
computed on the same GPU."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import opticalflow_ri_b200 as ofri  # noqa: E402
from opticalflow_ri_b200 import banded  # noqa: E402
from opticalflow_ri_b200.synthetic import synthetic_piv_pair  # noqa: E402


def tiled_rows(tile, r0, r1, size, dev):
    t = torch.from_numpy(tile).to(dev)
    reps = size // tile.shape[0]
    rows = torch.arange(r0, r1, device=dev) % tile.shape[0]
    return t[rows].repeat(1, reps).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--check-size", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--hs-niter", type=int, default=600)
    ap.add_argument("--ls-fuse", type=int, default=0)
    ap.add_argument("--hs-fuse-fast", type=int, default=0)
    ap.add_argument("--reserve-sms", type=int, default=-1, help="SMs an overlapped interior launch leaves to NCCL (-1 = default)")
    ap.add_argument("--exchange", type=int, default=0, help="HS sweeps between ghost-row exchanges (0 = default 32)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = ofri.Handle(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    h.set_stream(stream.cuda_stream)
    if args.exchange:
        h.set_option("band_exchange", args.exchange)
    if args.ls_fuse:
        h.set_option("ls_fuse", args.ls_fuse)
    if args.hs_fuse_fast:
        h.set_option("hs_fuse_fast", args.hs_fuse_fast)
    if args.reserve_sms >= 0:
        h.set_option("band_reserve_sms", args.reserve_sms)
    if world > 1:
        uid = [ofri.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init_nccl(rank, world, uid[0])
    mk = lambda n: ofri.make_params(ofri.hs_algo([45.0, 21.0], n), ofri.ls_algo(5.0, 60), filter_sigma=3.4,
                                    filter_opt_sigma=0.48, pyramid_levels=2, warping=True, bilinear=True, final_scaling=True)
    t0, t1 = synthetic_piv_pair(1024, 1024, 0)
    out = {"n_gpus": world}
    if args.check_size > 0:
        N = args.check_size
        p = mk(80)
        band = h.band_plan(N, N, p, rank, world)
        a = tiled_rows(t0, band.in0, band.in1, N, dev)
        b = tiled_rows(t1, band.in0, band.in1, N, dev)
        u, v = banded.flow_banded_rank(h, a, b, N, N, p, band)
        fa, fb = tiled_rows(t0, 0, N, N, dev), tiled_rows(t1, 0, N, N, dev)
        fu, fv = torch.empty_like(fa), torch.empty_like(fa)
        h.pyramidal_flow_ptr(fa.data_ptr(), fb.data_ptr(), 1, N, N, mk(80), fu.data_ptr(), fv.data_ptr(), None, device=True)
        torch.cuda.synchronize()
        bad = torch.tensor([int((u != fu[band.own0:band.own1]).sum().item()) + int((v != fv[band.own0:band.own1]).sum().item())],
                           device=dev)
        if world > 1:
            dist.all_reduce(bad)
        out["check"] = {"size": N, "mismatching_px_vs_single_gpu": int(bad.item())}
        del fa, fb, fu, fv, a, b, u, v
    N = args.size
    p = mk(args.hs_niter)
    band = h.band_plan(N, N, p, rank, world)
    a = tiled_rows(t0, band.in0, band.in1, N, dev)
    b = tiled_rows(t1, band.in0, band.in1, N, dev)
    times = []
    for rep in range(args.reps + 1):
        h.set_option("timing", 1 if rep == args.reps else 0)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        u, v = banded.flow_banded_rank(h, a, b, N, N, mk(args.hs_niter), band)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(t.item()))
    import statistics
    best = statistics.median(times)          # median of >= 5 repetitions (max over ranks each), not the minimum
    st = h.stage_timings()
    if rank == 0:
        px_it = 1.25 * N * N * (args.hs_niter + 60)
        out.update({"peer_allreduce": h.get_option("comm_peer_allreduce"), "reserve_sms": h.get_option("band_reserve_sms"), "size": N, "ms_per_pair": round(best, 2), "ms_per_pair_all": [round(x, 2) for x in times], "stat": "median", "pairs_per_s": round(1e3 / best, 4),
                    "gpix_iter_per_s": round(px_it / (best / 1e3) / 1e9, 1), "rows_owned": band.own1 - band.own0,
                    "rows_supplied": band.in1 - band.in0, "ghost": band.ghost, "exchange_every_sweeps": band.exchange,
                    "finite": bool(torch.isfinite(u).all().item()), "stages_rank0_ms": {k: round(x, 1) for k, x in st.items()}})
        print(json.dumps(out), flush=True)
    if world > 1:
        h.comm_destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
