#!/usr/bin/env python
"""bench.py -- headline benchmark of the HS + Liu-Shen pyramidal path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Headline workload (BASELINE.json configs[3], "batched synthetic PIV"): frame pairs of 1024 x 1024 float32, Horn-Schunck
(600 sweeps, alphas [21, 45]) + Liu-Shen (h = 5, 60 sweeps) with 2 pyramid levels, FILTER 3.4 / 3 taps, FILTER_OPT 0.48 /
5 taps (the parameters of examples/LiuSE_PyHSchunck_Fs3_4_PyrLvls2.py).  Pairs are independent: each rank (one process
per GPU) owns `--pairs-per-gpu` pairs (512 -> 4096 pairs on 8 GPUs), no data-path collective, weak scaling.  One "step"
= one pass of the whole path over the rank's pairs.

Printed JSON line (rank 0):
  value         pairs/s with inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e           the same through the host-pointer C-ABI call with PINNED host buffers (H2D + D2H inside the timed region)
  e2e_pageable  the same call with plain numpy (pageable) arrays -- what a drop-in caller passes; the library stages
                them through its pinned bounce ring
  roofline      the dominant kernel, live (stage timers = CUDA events around the kernel family's launches); traffic from
                the committed ncu captures (profiles/r2_traffic.json)
  configs       pairs/s of BASELINE configs 1 and 2 (Horn-Schunck only, 1 / 2 levels) on batched 512 x 512 pairs
  adapters      ms per pair of the Farneback and dense Lucas-Kanade adapters at the reference examples' parameters
  banded        BASELINE configs[4]: ONE 16384 x 16384 pair split into row bands over the N ranks (NCCL ghost-row
                exchange inside libofri.so); N = 1 runs the same entry point with one band.  Median ms per pair,
                Gpix-sweeps/s, a bit-for-bit check against the single-GPU path at 4096^2, max |d| against the
                REFERENCE's flow for the 2048^2 golden pair (tests/golden/big_2048.npz)
  cpu_baseline  the CPU oracle on a bounded sample (rank 0, N = 1)
`--impl reference` times the reference's own CPU implementation of the path (the unmodified reference modules when
/root/reference is importable, else the oracle port) on the box's host cores, on the same workload definition.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 1024
HS_NITER, HS_ALPHAS_IN_ORDER, LS_H, LS_ITERS, LEVELS = 600, [45.0, 21.0], 5.0, 60, 2
FILTER, FILTER_OPT = 3.4, 0.48
N_DISTINCT = 8            # distinct seeded synthetic pairs, tiled to fill the batch
BANDED_SIZE, BANDED_CHECK, BANDED_REPS = 16384, 4096, 5
REF_SRC = "/root/reference/src"


def pix_iters_per_pair(h=H, w=W):
    tot = 0
    scale = 1.0 / 2 ** (LEVELS - 1)
    for lv in range(1, LEVELS + 1):
        hl = h if lv == LEVELS else int(np.round(h * scale))
        wl = w if lv == LEVELS else int(np.round(w * scale))
        tot += hl * wl * (HS_NITER + LS_ITERS)
        scale *= 2
    return tot


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_json(name):
    p = os.path.join(ROOT, "profiles", name)
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(n_pairs, h=H, w=W):
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair
    base = [synthetic_piv_pair(h, w, seed=s) for s in range(min(N_DISTINCT, n_pairs))]
    a = np.empty((n_pairs, h, w), np.float32)
    b = np.empty((n_pairs, h, w), np.float32)
    for i in range(n_pairs):
        a[i], b[i] = base[i % len(base)]
    return a, b


def workload_config(args):
    """Identical for both arms (`--impl ours` and `--impl reference`): the workload, not the tuning."""
    return {"workload": "batched synthetic PIV (BASELINE configs[3]): %d pairs/GPU of %dx%d f32, HS 600 sweeps alphas "
                        "[21,45] + Liu-Shen h=5 60 sweeps, PyrLvls2, FILTER 3.4/3 taps, FILTER_OPT 0.48/5 taps"
                        % (args.pairs_per_gpu, H, W),
            "pairs_per_gpu": args.pairs_per_gpu, "H": H, "W": W, "parallelism": "pairs sharded by rank, no collective",
            "l2": "inputs larger than L2 (%.1f GiB of frames per step per GPU); no flush needed"
                  % (args.pairs_per_gpu * 2 * H * W * 4 / 2 ** 30)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms: the reference's own implementation (when importable) or the oracle port, on host cores
# ------------------------------------------------------------------------------------------------------------------
def reference_available():
    return os.path.isdir(REF_SRC) and os.path.exists(os.path.join(REF_SRC, "GenericPyramidalOpticalFlow.py"))


def _cpu_one(job):
    """One whole pair, single-threaded like the reference.  job = (seed, size, kind)."""
    seed, n, kind = job
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair
    i0, i1 = synthetic_piv_pair(n, n, seed=seed)
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import make_golden as MG                   # import_reference(): the unmodified modules + the matplotlib stub
        GPOF, HS, LS, GF, GKBE, WRAP = MG.import_reference()
        with MG.quiet():
            GPOF.genericPyramidalOpticalFlow(i0[:64, :64].copy(), i1[:64, :64].copy(), FILTER,
                                             HS.HSOpticalFlowAlgoAdapter([21, 45], 2), LEVELS, 1, FILTER_OPT,
                                             LS.LiuShenOpticalFlowAlgoAdapter(LS_H))      # numba JIT warm-up
            t = time.perf_counter()
            GPOF.genericPyramidalOpticalFlow(i0, i1, FILTER, HS.HSOpticalFlowAlgoAdapter([21, 45], HS_NITER), LEVELS, 1,
                                             FILTER_OPT, LS.LiuShenOpticalFlowAlgoAdapter(LS_H))
            return time.perf_counter() - t
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ofri_oracle as O
    t = time.perf_counter()
    O.pyramidal_flow(i0, i1, FILTER, O.HSParams([21, 45], HS_NITER), LEVELS, 1, FILTER_OPT, O.LSParams(LS_H))
    return time.perf_counter() - t


def cpu_step(cores, size, kind, seed0=0):
    """`cores` processes, each running ONE size^2 pair with the full parameters.  Returns (pairs/s at 1024^2, description)."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    jobs = [(seed0 + s, size, kind) for s in range(cores)]
    if cores == 1:
        times = [_cpu_one(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            times = pool.map(_cpu_one, jobs)
    wall = time.perf_counter() - t0
    scale = (H * W) / float(size * size)
    per_pair = statistics.mean(times) * scale
    desc = "%d process(es) x 1 synthetic pair of %dx%d px, full parameters (HS 600 + LS 60 sweeps, 2 levels), mean %.1f " \
           "s/pair%s; wall %.1f s" % (cores, size, size, statistics.mean(times),
                                      "" if size == H else ", scaled x%.2f by pixel count to 1024x1024" % scale, wall)
    return cores / per_pair, desc


def cpu_baseline(cores, size, kind="port"):
    v, desc = cpu_step(cores, size, kind)
    return {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": desc,
            "gpix_iter_per_s": v * pix_iters_per_pair() / 1e9}


def run_reference(args):
    """The reference arm: full 1024 x 1024 pairs (no pixel scaling), one per host core per step.  A step costs 30-70 s of
    wall time on the box (all cores busy), so the number of timed steps is bounded by --ref-budget seconds (stated in the
    line: steps = timed steps actually run, steps_requested = the driver's K)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = args.cpu_cores or os.cpu_count() or 1
    kind = "reference" if (reference_available() and not args.ref_port) else "port"
    size = args.cpu_sample or H
    t_all = time.perf_counter()
    # warm-up: imports, page cache, (reference: numba compilation happens inside every worker before its timer starts)
    warm_done = 0
    for _ in range(min(args.warmup, 1)):
        cpu_step(min(cores, 2), 96, kind)
        warm_done += 1
    vals, last_desc = [], ""
    steps_req = max(1, args.steps)
    for i in range(steps_req):
        if i > 0 and time.perf_counter() - t_all > args.ref_budget:
            break
        v, last_desc = cpu_step(cores, size, kind, seed0=i * cores)
        vals.append(v)
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": "frame_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": warm_done, "steps_requested": steps_req, "warmup_requested": args.warmup,
            "steps_note": "a step = one full 1024x1024 pair per host core (tens of seconds of wall time with every core "
                          "busy); timed steps are bounded by --ref-budget = %d s so the arm ends within minutes; warm-up = "
                          "one 96x96 pair on 2 processes (imports; the reference's numba JIT is warmed inside every "
                          "worker before its timer starts)" % args.ref_budget,
            "ms_per_step": 1e3 * cores / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gpix_iter_per_s": v * pix_iters_per_pair() / 1e9,
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": kind, "sample": last_desc,
                             "gpix_iter_per_s": v * pix_iters_per_pair() / 1e9},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# row-band mode: ONE very large pair over the ranks (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------------------------
def tiled_rows(torch, tile, r0, r1, size, dev):
    t = torch.from_numpy(tile).to(dev)
    reps = size // tile.shape[0]
    rows = torch.arange(r0, r1, device=dev) % tile.shape[0]
    return t[rows].repeat(1, reps).contiguous()


def run_banded(args, torch, dist, ofri, h, rank, world, dev):
    from opticalflow_ri_b200 import banded
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair

    def mk():
        return ofri.make_params(ofri.hs_algo(HS_ALPHAS_IN_ORDER, HS_NITER), ofri.ls_algo(LS_H, LS_ITERS),
                                filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT, pyramid_levels=LEVELS, warping=True,
                                bilinear=True, final_scaling=True)

    def allmax(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    out = {"n_gpus": world, "backend": "single band (no communicator)"}
    if world > 1:
        uid = [ofri.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        h.comm_init_nccl(rank, world, uid[0])
        out["backend"] = "inside libofri.so: ncclSend/ncclRecv ghost rows; scalar all-reduces " + (
            "by a one-kernel reduction over NVLink peer memory (CUDA IPC mailboxes)" if h.get_option("comm_peer_allreduce")
            else "by ncclAllReduce")
    t0, t1 = synthetic_piv_pair(1024, 1024, 0)
    # (1) the REFERENCE's flow for the 2048^2 golden pair (tests/golden/big_2048.npz), banded over the ranks
    gpath = os.path.join(ROOT, "tests", "golden", "big_2048.npz")
    if os.path.exists(gpath):
        g = np.load(gpath)
        N = 2048
        p = mk()
        band = h.band_plan(N, N, p, rank, world)
        a = torch.from_numpy(g["im0"][band.in0:band.in1].astype(np.float32)).to(dev)
        b = torch.from_numpy(g["im1"][band.in0:band.in1].astype(np.float32)).to(dev)
        u, v = banded.flow_banded_rank(h, a, b, N, N, p, band)
        gu = torch.from_numpy(g["U"][band.own0:band.own1]).to(dev)
        gv = torch.from_numpy(g["V"][band.own0:band.own1]).to(dev)
        out["max_abs_vs_reference_golden_2048"] = allmax(max(float((u - gu).abs().max()), float((v - gv).abs().max())))
        out["golden"] = "tests/golden/big_2048.npz: the unmodified reference's flow for a seeded synthetic 2048x2048 " \
                        "pair, full parameters (oracle/make_golden.py --big 2048); tolerance 1e-4 px"
        del a, b, u, v, gu, gv
    # (2) bit-for-bit against the single-GPU (batched) path at BANDED_CHECK^2, full parameters
    N = args.banded_check
    if N > 0:
        p = mk()
        band = h.band_plan(N, N, p, rank, world)
        a = tiled_rows(torch, t0, band.in0, band.in1, N, dev)
        b = tiled_rows(torch, t1, band.in0, band.in1, N, dev)
        u, v = banded.flow_banded_rank(h, a, b, N, N, p, band)
        fa, fb = tiled_rows(torch, t0, 0, N, N, dev), tiled_rows(torch, t1, 0, N, N, dev)
        fu, fv = torch.empty_like(fa), torch.empty_like(fa)
        h.pyramidal_flow_ptr(fa.data_ptr(), fb.data_ptr(), 1, N, N, mk(), fu.data_ptr(), fv.data_ptr(), None, device=True)
        torch.cuda.synchronize()
        bad = int((u != fu[band.own0:band.own1]).sum().item()) + int((v != fv[band.own0:band.own1]).sum().item())
        out["check_size"] = N
        out["mismatching_px_vs_single_gpu"] = int(allsum(bad))
        del fa, fb, fu, fv, a, b, u, v
        torch.cuda.empty_cache()
    # (3) timing at BANDED_SIZE^2
    N = args.banded_size
    p = mk()
    band = h.band_plan(N, N, p, rank, world)
    a = tiled_rows(torch, t0, band.in0, band.in1, N, dev)
    b = tiled_rows(torch, t1, band.in0, band.in1, N, dev)
    times = []
    for rep in range(args.banded_reps + 2):          # two untimed repetitions first (the first one allocates the workspace)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        u, v = banded.flow_banded_rank(h, a, b, N, N, mk(), band)
        e1.record()
        torch.cuda.synchronize()
        t = allmax(e0.elapsed_time(e1))
        if rep > 1:
            times.append(t)
    h.set_option("timing", 1)
    u, v = banded.flow_banded_rank(h, a, b, N, N, mk(), band)
    torch.cuda.synchronize()
    st = h.stage_timings()
    h.set_option("timing", 0)
    finite = bool(torch.isfinite(u).all().item())
    ms = statistics.median(times)
    px_it = 1.25 * N * N * (HS_NITER + LS_ITERS)
    n1 = committed_json("r2_banded_n1.json")
    out.update({"size": N, "ms_per_pair": ms, "ms_per_pair_all": [round(x, 2) for x in times], "stat": "median of %d (max over ranks each)" % len(times),
                "pairs_per_s": 1e3 / ms, "gpix_iter_per_s": px_it / (ms / 1e3) / 1e9,
                "rows_owned": band.own1 - band.own0, "rows_supplied": band.in1 - band.in0, "ghost_rows": band.ghost,
                "exchange_every_sweeps": band.exchange, "reserve_sms": h.get_option("band_reserve_sms") if world > 1 else 0,
                "finite": finite, "stages_rank0_ms": {k: round(x, 2) for k, x in st.items()}})
    if world == 1:
        out["efficiency_vs_n1_ms"] = 1.0
        out["n1_ms_per_pair"] = ms
    elif n1 and n1.get("size") == N:
        out["n1_ms_per_pair"] = n1["ms_per_pair"]
        out["n1_source"] = "profiles/r2_banded_n1.json (N = 1 run of this bench at %s)" % n1.get("head", "?")
        out["efficiency_vs_n1_ms"] = n1["ms_per_pair"] / (world * ms)
    out["limiter"] = ("per rank and solve: 600/32 grouped ncclSend/ncclRecv ghost-row exchanges of 32 rows x W x (U,V) "
                      "overlapped with the interior tiles (the overlapped launches leave band_reserve_sms SMs to the NCCL "
                      "kernel), 15 all-reduces of the Liu-Shen residual sums per level (stream-ordered, one per fused block "
                      "of 4 sweeps; %s), 2 x ghost_rows redundant rows per band and level, fixed per-launch cost on bands "
                      "1/N as tall" % ("one 1-CTA kernel each over NVLink peer memory" if world > 1 and
                                       h.get_option("comm_peer_allreduce") else "ncclAllReduce"))
    if world > 1:
        h.comm_destroy()
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import opticalflow_ri_b200 as ofri

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = ofri.Handle(local)
    tuning = {}
    for k, v in (("hs_fuse", args.hs_fuse), ("hs_variant", args.hs_variant), ("ls_fuse", args.ls_fuse),
                 ("ls_variant", args.ls_variant), ("hs_precise", args.hs_precise), ("chunk_pairs", args.chunk_pairs)):
        if v is not None:
            h.set_option(k, v)
    P = args.pairs_per_gpu

    def mk(levels=LEVELS, with_ls=True, alphas=HS_ALPHAS_IN_ORDER):
        return ofri.make_params(ofri.hs_algo(alphas, HS_NITER), ofri.ls_algo(LS_H, LS_ITERS) if with_ls else None,
                                filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT if with_ls else None,
                                pyramid_levels=levels, warping=True, bilinear=True, final_scaling=True)

    params = mk()
    a_np, b_np = make_inputs(P)
    # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event times it.  (torch's default
    # stream has handle 0, which ofri_set_stream treats as "use the handle's own stream".)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    h.set_stream(stream.cuda_stream)
    d_a = torch.from_numpy(a_np).cuda()
    d_b = torch.from_numpy(b_np).cuda()
    d_u = torch.empty_like(d_a)
    d_v = torch.empty_like(d_a)

    def step_dev():
        h.pyramidal_flow_ptr(d_a.data_ptr(), d_b.data_ptr(), P, H, W, params, d_u.data_ptr(), d_v.data_ptr(), None,
                             device=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    n0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    ms_max = allmax(e0.elapsed_time(e1))
    launches = h.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    value = world * P * args.steps / (ms_max / 1e3)
    for k in ("hs_fuse", "hs_variant", "hs_precise", "ls_fuse", "ls_variant", "auto_fuse", "last_hs_fuse_fine",
              "last_hs_fuse_coarse", "last_ls_fuse", "last_chunk_pairs"):
        tuning[k] = h.get_option(k)

    # ---- roofline of the dominant kernel, live: one instrumented step with per-stage CUDA events -------------------
    # Stage timers bracket exactly the launches of one kernel family on the launching stream.  Algorithmic bytes
    # (DESIGN.md section 4): fused Horn-Schunck sweeps 28 B / pixel / launch, fused Liu-Shen sweeps 48 B / pixel / launch.
    roof = None
    stage_ms = {}
    if rank == 0:
        h.set_option("timing", 1)
        step_dev()
        torch.cuda.synchronize()
        h.synchronize()                      # also collects the per-stage CUDA-event timings
        stage_ms = h.stage_timings()
        h.set_option("timing", 0)
        Tf, Tc = max(tuning["last_hs_fuse_fine"], 1), max(tuning["last_hs_fuse_coarse"], 1)
        Tl = max(tuning["last_ls_fuse"], 1)
        peak, peak_src = measured_peak()
        px_fine, px_coarse = H * W, int(np.round(H * 0.5)) * int(np.round(W * 0.5))
        chunk = tuning["last_chunk_pairs"] or min(P, 64)   # pairs per launch (the library's chunking)
        nchunks = -(-P // chunk)
        traffic_db = committed_json("r2_traffic.json") or {}

        def entry(key, kernel, stage, bytes_px, px_levels, T, sweeps):
            ms = stage_ms.get(stage)
            if not ms:
                return None
            per_level = -(-sweeps // max(T, 1))
            algo = sum(bytes_px * px * P * per_level for px in px_levels)
            ach = algo / (ms / 1e3) / 1e9
            nlaunch = per_level * len(px_levels) * nchunks
            tr = traffic_db.get(key)
            traffic = None
            if tr and tr.get("pairs"):
                traffic = int((tr["dram_read_bytes"] + tr["dram_write_bytes"]) * chunk / float(tr["pairs"]))
            return {"key": key, "kernel": kernel, "stage": stage, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "frac_of_nominal_8000_GBs": ach / 8000.0, "fused_sweeps": T,
                    "stage_ms": ms, "launches": nlaunch, "avg_launch_ms": ms / nlaunch,
                    "algorithmic_bytes_per_launch": algo / nlaunch,
                    "traffic": traffic, "traffic_source": (tr or {}).get("source"),
                    "gpix_sweeps_per_s": sum(px_levels) * P * sweeps / (ms / 1e3) / 1e9,
                    "frac_in_unfused_bytes": sum(px_levels) * P * sweeps * bytes_px / (ms / 1e3) / 1e9 / peak}

        # Which Horn-Schunck kernels ran is read off the stage timers: "hs_iterate" = fast kernel on the finest level,
        # "hs_iterate_coarse" = fast kernel on the coarser level (hs_precise = 0, or 1 with the Liu-Shen refinement after
        # it: this workload), "hs_iterate_precise" = reference-arithmetic kernel.
        kernels = [k for k in (
            entry("hs_tma_fast_fine", "hs_tma_kernel<T=%d,R=8,NRG=8,fast> (persistent TMA-fed fused Horn-Schunck sweeps, "
                  "finest level)" % Tf, "hs_iterate", 28.0, [px_fine], Tf, HS_NITER),
            entry("hs_tma_fast_coarse", "hs_tma_kernel<T=%d,R=8,NRG=8,fast> (same kernel, coarse level 512x512)" % Tc,
                  "hs_iterate_coarse", 28.0, [px_coarse], Tc, HS_NITER),
            entry("hs_tma_precise", "hs_tma_kernel<T=%d,R=4,NRG=8,precise> (same, reference arithmetic)" % Tc,
                  "hs_iterate_precise", 28.0, [px_coarse] if stage_ms.get("hs_iterate") else [px_coarse, px_fine], Tc,
                  HS_NITER),
            entry("ls_tma", "ls_tma_kernel<T=%d,R=4,NRG=8> (persistent TMA-fed fused Liu-Shen sweeps, both levels)" % Tl,
                  "ls_iterate", 48.0, [px_coarse, px_fine], Tl, LS_ITERS)) if k]
        if kernels:
            dom = max(kernels, key=lambda k: k["stage_ms"])
            roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "frac_of_nominal_8000_GBs": dom["frac_of_nominal_8000_GBs"],
                    "traffic": dom["traffic"], "traffic_source": dom["traffic_source"], "peak_source": peak_src,
                    "pairs_per_launch": chunk,
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                    "avg_launch_ms": dom["avg_launch_ms"], "share_of_step": dom["stage_ms"] / sum(stage_ms.values()),
                    "note": "achieved = algorithmic bytes of the kernel's launches in one step / their CUDA-event time "
                            "(stage timer on the launching stream); 28 B/px/launch = read U,V,a,b,c + write U,V, for T "
                            "fused sweeps (48 B/px/launch for Liu-Shen: u,v + 8 coefficient planes); frac_in_unfused_bytes "
                            "= the same sweeps/s expressed in the bytes an unfused (T=1) sweep moves; traffic = dram "
                            "read+write bytes of one launch from the committed ncu capture named in traffic_source, "
                            "scaled to pairs_per_launch (null when no capture of that kernel is committed)",
                    "kernels": kernels}
    # ---- end-to-end through the host-pointer C-ABI call: pinned host buffers, then pageable (numpy) ones -------------------
    Pe = min(P, args.e2e_pairs)
    ha = torch.from_numpy(a_np[:Pe]).pin_memory()
    hb = torch.from_numpy(b_np[:Pe]).pin_memory()
    hu = torch.empty_like(ha).pin_memory()
    hv = torch.empty_like(ha).pin_memory()
    h.set_stream(0)

    def step_e2e():
        h.pyramidal_flow_ptr(ha.data_ptr(), hb.data_ptr(), Pe, H, W, params, hu.data_ptr(), hv.data_ptr(), None,
                             device=False)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_e2e()          # synchronous: returns when the last D2H has landed
    torch.cuda.synchronize()
    e2e_val = world * Pe * args.e2e_steps / allmax(time.perf_counter() - t0)
    check = float(hu[0].abs().max())
    del ha, hb, hu, hv
    # pageable: the numpy arrays themselves (all P pairs), outputs into fresh numpy arrays
    pu = np.empty_like(a_np)
    pv = np.empty_like(a_np)

    def step_pageable():
        h.pyramidal_flow_ptr(a_np.ctypes.data, b_np.ctypes.data, P, H, W, params, pu.ctypes.data, pv.ctypes.data, None,
                             device=False)

    step_pageable()
    host_path = h.get_option("last_host_path")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step_pageable()
    torch.cuda.synchronize()
    e2e_pageable = world * P * args.e2e_steps / allmax(time.perf_counter() - t0)
    check_p = float(np.abs(pu[0]).max())
    del pu, pv

    # ---- BASELINE configs 1 and 2 (Horn-Schunck only) on batched 512 x 512 pairs ------------------------------------------
    configs = None
    if not args.no_configs:
        h.set_stream(stream.cuda_stream)
        Pc, Hc = args.configs_pairs, 512
        ca, cb = make_inputs(Pc, Hc, Hc)
        dca, dcb = torch.from_numpy(ca).cuda(), torch.from_numpy(cb).cuda()
        dcu, dcv = torch.empty_like(dca), torch.empty_like(dca)
        configs = {"pairs_per_gpu": Pc, "H": Hc, "W": Hc, "unit": "pairs/s (device-resident, CUDA events, max over ranks)"}
        for name, pr in (("config1_HS_Fs3_4", mk(1, False, [21.0])), ("config2_HS_Fs3_4_PyrLvls2", mk(2, False))):
            def run_c():
                h.pyramidal_flow_ptr(dca.data_ptr(), dcb.data_ptr(), Pc, Hc, Hc, pr, dcu.data_ptr(), dcv.data_ptr(), None,
                                     device=True)
            run_c()
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(2):
                run_c()
            c1.record(stream)
            barrier()
            msc = allmax(c0.elapsed_time(c1))
            px_it = sum(int(np.round(Hc * s)) ** 2 for s in ([1.0] if "config1" in name else [0.5, 1.0])) * HS_NITER
            configs[name] = {"pairs_per_s": world * Pc * 2 / (msc / 1e3),
                             "gpix_sweeps_per_s": world * Pc * 2 * px_it / (msc / 1e3) / 1e9}
        del dca, dcb, dcu, dcv
    # ---- the reference's two other adapters at their example parameters, batched 512 x 512 pairs ---------------
    adapters = None
    if not args.no_adapters:
        sys.path.insert(0, ofri.SRC_DIR)
        try:
            from Farneback_PyCL import Farneback_PyCL
            from denseLucasKanade_PyCL import denseLucasKanade_PyCl
        finally:
            sys.path.remove(ofri.SRC_DIR)
        h.set_stream(stream.cuda_stream)
        Pa, Ha = 64, 512
        aa, ab = make_inputs(Pa, Ha, Ha)
        daa, dab = torch.from_numpy(aa).cuda(), torch.from_numpy(ab).cuda()
        dau, dav = torch.empty_like(daa), torch.empty_like(daa)
        h.set_farneback(Farneback_PyCL().native_params())
        h.set_lk(denseLucasKanade_PyCl(Niter=5, halfWindow=13).native_params())
        adapters = {"pairs": Pa, "H": Ha, "W": Ha, "unit": "ms per pair (device-resident, CUDA events, one-level driver call, "
                                                          "adapter as main, best of 4)"}
        for name, algo, sigma in (("farneback_w33_it5_poly7", ofri.fb_algo(), 0.0), ("dense_lk_hw13_it5", ofri.lk_algo(), 2.0)):
            pr = ofri.make_params(algo, None, filter_sigma=sigma, pyramid_levels=1, k_levels=1, warping=False,
                                  final_scaling=False)
            best = None
            for _ in range(4):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                h.pyramidal_flow_ptr(daa.data_ptr(), dab.data_ptr(), Pa, Ha, Ha, pr, dau.data_ptr(), dav.data_ptr(), None,
                                     device=True)
                a1.record(stream)
                a1.synchronize()
                ms = a0.elapsed_time(a1) / Pa
                best = ms if best is None or ms < best else best
            adapters[name] = {"ms_per_pair": best, "mpx_per_s": Ha * Ha / best / 1e3}
        del daa, dab, dau, dav
    # ---- row-band mode (BASELINE configs[4]) ---------------------------------------------------------------------------------
    banded_out = None
    if not args.no_banded:
        del d_a, d_b, d_u, d_v
        torch.cuda.empty_cache()
        h.set_stream(stream.cuda_stream)
        try:
            banded_out = run_banded(args, torch, dist, ofri, h, rank, world, dev)
        except Exception as ex:           # reported in the line, never silently dropped
            banded_out = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if rank == 0:
        cpu = cpu_baseline(1, args.cpu_sample or 768) if (world == 1 and not args.no_cpu) else None
        line = {"metric": "frame_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (%d seeded PIV pairs tiled)"
                % min(N_DISTINCT, P),
                "gpix_iter_per_s": value * pix_iters_per_pair() / 1e9,
                "config": workload_config(args), "tuning": tuning,
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": int(Pe * 2 * H * W * 4),
                        "d2h_bytes_per_step": int(Pe * 2 * H * W * 4), "pairs_per_step": Pe, "steps": args.e2e_steps,
                        "host_memory": "pinned", "result_check_max_abs_u": check},
                "e2e_pageable": {"value": e2e_pageable, "unit": "pairs/s", "h2d_bytes_per_step": int(P * 2 * H * W * 4),
                                 "d2h_bytes_per_step": int(P * 2 * H * W * 4), "pairs_per_step": P, "steps": args.e2e_steps,
                                 "host_memory": "pageable numpy arrays (library path %d: %s)"
                                 % (host_path, "pinned bounce ring + 2 host threads" if host_path == 2 else "direct copies"),
                                 "ratio_to_pinned": e2e_pageable / e2e_val, "result_check_max_abs_u": check_p},
                "configs": configs, "adapters": adapters, "banded": banded_out,
                "gpu_launches": int(launches), "clocks": clocks,
                "stage_ms_one_step": stage_ms}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=512)
    ap.add_argument("--e2e-pairs", type=int, default=256)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="side of the sample pair the CPU arms time (0 = default: reference arm 1024 = the workload's "
                         "frame, no scaling; cpu_baseline of the GPU arm 768, ~12 s of single-core work)")
    ap.add_argument("--cpu-cores", type=int, default=0, help="host processes of the reference arm (0 = all cores)")
    ap.add_argument("--ref-budget", type=int, default=150, help="reference arm: no new timed step after this many seconds")
    ap.add_argument("--ref-port", action="store_true", help="reference arm: use the oracle port even if /root/reference exists")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-banded", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-adapters", action="store_true")
    ap.add_argument("--banded-size", type=int, default=BANDED_SIZE)
    ap.add_argument("--banded-check", type=int, default=BANDED_CHECK)
    ap.add_argument("--banded-reps", type=int, default=BANDED_REPS)
    ap.add_argument("--configs-pairs", type=int, default=256)
    ap.add_argument("--hs-fuse", type=int, default=None)
    ap.add_argument("--hs-variant", type=int, default=None)
    ap.add_argument("--ls-fuse", type=int, default=None)
    ap.add_argument("--ls-variant", type=int, default=None)
    ap.add_argument("--hs-precise", type=int, default=None)
    ap.add_argument("--chunk-pairs", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
