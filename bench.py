#!/usr/bin/env python
"""bench.py -- headline benchmark of the HS + Liu-Shen pyramidal path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Workload (BASELINE.json configs[3], "batched synthetic PIV"): frame pairs of 1024 x 1024 float32, Horn-Schunck
(600 sweeps, alphas [21, 45]) + Liu-Shen (h = 5, 60 sweeps) with 2 pyramid levels, FILTER 3.4 / 3 taps, FILTER_OPT 0.48 /
5 taps (the parameters of examples/LiuSE_PyHSchunck_Fs3_4_PyrLvls2.py).  Pairs are independent: each rank (one process
per GPU) owns `--pairs-per-gpu` pairs (512 -> 4096 pairs on 8 GPUs), no data-path collective, weak scaling.  One "step"
= one pass of the whole path over the rank's pairs.

Printed JSON line (rank 0): value = pairs/s with inputs resident in HBM (CUDA events, max over ranks); e2e = the same
through the host-pointer C-ABI call with pinned host buffers (H2D + D2H inside the timed region); roofline = the
dominant kernel (fused HS sweeps) measured live with CUDA events; cpu_baseline = the oracle port on a bounded sample.
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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel (64 pairs of 1024 x 1024, T = 4) from
# the committed `ncu --set full` capture profiles/r1_ncu_hs_tma_fast_64pairs.txt; None until that capture exists
ROOFLINE_TRAFFIC_BYTES_PER_LAUNCH = 1342233000 + 507894528   # read + write, 396 us launch


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


def make_inputs(n_pairs):
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair
    base = [synthetic_piv_pair(H, W, seed=s) for s in range(min(N_DISTINCT, n_pairs))]
    a = np.empty((n_pairs, H, W), np.float32)
    b = np.empty((n_pairs, H, W), np.float32)
    for i in range(n_pairs):
        a[i], b[i] = base[i % len(base)]
    return a, b


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on host cores (bounded sample)
# ------------------------------------------------------------------------------------------------------------------
def _cpu_one(seed_and_size):
    seed, n = seed_and_size
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ofri_oracle as O
    from opticalflow_ri_b200.synthetic import synthetic_piv_pair
    i0, i1 = synthetic_piv_pair(n, n, seed=seed)
    t = time.perf_counter()
    O.pyramidal_flow(i0, i1, FILTER, O.HSParams([21, 45], HS_NITER), LEVELS, 1, FILTER_OPT, O.LSParams(LS_H))
    return time.perf_counter() - t


def cpu_baseline(cores, sample_px=768):
    """Oracle (kind 'port': numpy restatement of the reference, single-threaded per pair like the reference) on
    `cores` processes, each running ONE sample_px^2 pair with the full parameters; scaled to 1024^2 pairs by pixel count
    (the path is linear in pixels; stated in `sample`)."""
    import multiprocessing as mp
    t0 = time.perf_counter()
    if cores == 1:
        times = [_cpu_one((0, sample_px))]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            times = pool.map(_cpu_one, [(s, sample_px) for s in range(cores)])
    wall = time.perf_counter() - t0
    per_pair_1024 = statistics.mean(times) * (H * W) / float(sample_px * sample_px)
    pairs_s = cores / per_pair_1024
    return {"value": pairs_s, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": "%d process(es) x 1 synthetic pair of %dx%d px, full parameters (HS 600 + LS 60 sweeps, 2 levels), "
                      "mean %.1f s/pair, scaled x%.2f by pixel count to 1024x1024; wall %.1f s"
                      % (cores, sample_px, sample_px, statistics.mean(times), (H * W) / float(sample_px ** 2), wall),
            "gpix_iter_per_s": pairs_s * pix_iters_per_pair() / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = args.cpu_cores or os.cpu_count() or 1
    steps = max(1, args.steps)
    vals = []
    last = None
    t_all = time.perf_counter()
    for i in range(args.warmup + steps):
        if i > 0 and time.perf_counter() - t_all > 150:     # keep the whole run within a few minutes
            break
        last = cpu_baseline(cores, sample_px=args.cpu_sample)
        if i >= args.warmup:
            vals.append(last["value"])
    if not vals:
        vals = [last["value"]]
    v = statistics.mean(vals)
    line = {"impl": "reference", "metric": "frame_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * cores / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gpix_iter_per_s": v * pix_iters_per_pair() / 1e9,
            "config": workload_config(args, None),
            "cpu_baseline": dict(last, value=v),
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, handle):
    cfg = {"workload": "batched synthetic PIV (BASELINE configs[3]): %d pairs/GPU of %dx%d f32, HS 600 sweeps alphas "
                       "[21,45] + Liu-Shen h=5 60 sweeps, PyrLvls2, FILTER 3.4/3 taps, FILTER_OPT 0.48/5 taps"
                       % (args.pairs_per_gpu, H, W),
           "pairs_per_gpu": args.pairs_per_gpu, "H": H, "W": W, "parallelism": "pairs sharded by rank, no collective",
           "l2": "inputs larger than L2 (%.1f GiB of frames per step per GPU); no flush needed"
                 % (args.pairs_per_gpu * 2 * H * W * 4 / 2 ** 30)}
    if handle is not None:
        cfg.update({k: handle.get_option(k) for k in ("hs_fuse", "hs_variant", "hs_precise", "ls_fuse", "ls_variant")})
    return cfg


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h = ofri.Handle(local)
    for k, v in (("hs_fuse", args.hs_fuse), ("hs_variant", args.hs_variant), ("ls_fuse", args.ls_fuse),
                 ("ls_variant", args.ls_variant), ("hs_precise", args.hs_precise), ("chunk_pairs", args.chunk_pairs)):
        if v is not None:
            h.set_option(k, v)
    P = args.pairs_per_gpu
    params = ofri.make_params(ofri.hs_algo(HS_ALPHAS_IN_ORDER, HS_NITER), ofri.ls_algo(LS_H, LS_ITERS),
                              filter_sigma=FILTER, filter_opt_sigma=FILTER_OPT, pyramid_levels=LEVELS, warping=True,
                              bilinear=True, final_scaling=True)
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
    ms = e0.elapsed_time(e1)
    launches = h.launch_count - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * P * args.steps / (ms_max / 1e3)

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
        T = h.get_option("hs_fuse")
        Tl = max(h.get_option("ls_fuse"), 1)
        peak, peak_src = measured_peak()
        px_fine, px_coarse = H * W, int(np.round(H * 0.5)) * int(np.round(W * 0.5))
        nl = -(-HS_NITER // max(T, 1))            # launches per level
        chunk = h.get_option("last_chunk_pairs") or min(P, 64)   # pairs per launch (the library's chunking)
        nchunks = -(-P // chunk)

        def entry(kernel, stage, bytes_px, px_levels, launches_per_level, sweeps):
            ms = stage_ms.get(stage)
            if not ms:
                return None
            algo = sum(bytes_px * px * P * launches_per_level for px in px_levels)
            ach = algo / (ms / 1e3) / 1e9
            nlaunch = launches_per_level * len(px_levels) * nchunks
            return {"kernel": kernel, "stage": stage, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "frac_of_nominal_8000_GBs": ach / 8000.0,
                    "stage_ms": ms, "launches": nlaunch, "avg_launch_ms": ms / nlaunch,
                    "algorithmic_bytes_per_launch": algo / nlaunch,
                    "gpix_sweeps_per_s": sum(px_levels) * P * sweeps / (ms / 1e3) / 1e9,
                    "frac_in_unfused_bytes": sum(px_levels) * P * sweeps * bytes_px / (ms / 1e3) / 1e9 / peak}

        # Which Horn-Schunck kernels ran is read off the stage timers: "hs_iterate" = fast kernel on the finest level,
        # "hs_iterate_coarse" = fast kernel on the coarser level (hs_precise = 0, or 1 with the Liu-Shen refinement after
        # it: this workload), "hs_iterate_precise" = reference-arithmetic kernel.
        kernels = [k for k in (
            entry("hs_tma_kernel<T=%d,R=8,NRG=8,fast> (persistent TMA-fed fused Horn-Schunck sweeps, finest level)" % T,
                  "hs_iterate", 28.0, [px_fine], nl, HS_NITER),
            entry("hs_tma_kernel<T=%d,R=8,NRG=8,fast> (same kernel, coarse level 512x512: 15 %% of the tile area lies "
                  "beyond the image border)" % T, "hs_iterate_coarse", 28.0, [px_coarse], nl, HS_NITER),
            entry("hs_tma_kernel<T=%d,R=4,NRG=8,precise> (same, reference arithmetic)" % T,
                  "hs_iterate_precise", 28.0, [px_coarse] if stage_ms.get("hs_iterate") else [px_coarse, px_fine], nl,
                  HS_NITER),
            entry("ls_tma_kernel<T=%d,R=4,NRG=8> (persistent TMA-fed fused Liu-Shen sweeps, both levels)" % Tl,
                  "ls_iterate", 48.0, [px_coarse, px_fine], -(-LS_ITERS // Tl), LS_ITERS)) if k]
        if kernels:
            dom = max(kernels, key=lambda k: k["stage_ms"])
            roof = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "frac_of_nominal_8000_GBs": dom["frac_of_nominal_8000_GBs"], "traffic": int(ROOFLINE_TRAFFIC_BYTES_PER_LAUNCH * chunk / 64.0), "peak_source": peak_src,
                    "pairs_per_launch": chunk,
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                    "avg_launch_ms": dom["avg_launch_ms"], "share_of_step": dom["stage_ms"] / sum(stage_ms.values()),
                    "note": "achieved = algorithmic bytes of the kernel's launches in one step / their CUDA-event time "
                            "(stage timer on the launching stream); 28 B/px/launch = read U,V,a,b,c + write U,V, for T "
                            "fused sweeps; frac_in_unfused_bytes = the same sweeps/s expressed in the 28 B/px/sweep an "
                            "unfused (T=1) sweep moves; traffic = dram read+write bytes per launch (captured for a 64-pair launch: 1.850 GB vs 1.879 GB algorithmic, scaled to pairs_per_launch) from the committed "
                            "ncu capture (profiles/), null until captured for this launch shape",
                    "kernels": kernels}
    # ---- end-to-end through the host-pointer C-ABI call (pinned host buffers) -----------------------------------------
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
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * Pe * args.e2e_steps / float(te.item())
    check = float(hu[0].abs().max())

    if rank == 0:
        cpu = cpu_baseline(1, sample_px=args.cpu_sample) if (world == 1 and not args.no_cpu) else None
        line = {"metric": "frame_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (%d seeded PIV pairs tiled)"
                % min(N_DISTINCT, P),
                "gpix_iter_per_s": value * pix_iters_per_pair() / 1e9,
                "config": workload_config(args, h),
                "roofline": roof, "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "pairs/s", "h2d_bytes_per_step": int(Pe * 2 * H * W * 4),
                        "d2h_bytes_per_step": int(Pe * 2 * H * W * 4), "pairs_per_step": Pe, "steps": args.e2e_steps,
                        "host_memory": "pinned", "result_check_max_abs_u": check},
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
    ap.add_argument("--cpu-sample", type=int, default=768,
                    help="side of the sample pair the CPU arms time (768: ~12 s of single-core work per pair)")
    ap.add_argument("--cpu-cores", type=int, default=0, help="host processes of the reference arm (0 = all cores)")
    ap.add_argument("--no-cpu", action="store_true")
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
