"""N > 1 host logic on CPU (gloo, world_size 2 and 3): the row-band protocol with real message passing against the
oracle, and the launch contract of `bench.py --impl reference` under torchrun (rank 0 alone works and prints)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(free_port()), script] + list(args)
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("world", [2, 3])
def test_band_protocol_gloo(world):
    r = torchrun(world, os.path.join("tests", "mp_band_worker.py"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MP_BAND_RESULT")]
    assert len(line) == 1 and "mismatching_px=0" in line[0] and "ranks=%d" % world in line[0], r.stdout[-2000:]


def test_bench_reference_arm_under_torchrun():
    r = torchrun(2, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--cpu-sample", "40",
                 "--cpu-cores", "2", "--ref-port")
    assert r.returncode == 0, r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["steps"] == 2 and d["steps_requested"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert set(d["config"]) == {"workload", "pairs_per_gpu", "H", "W", "parallelism", "l2"}     # no tuning keys: same as the GPU arm's


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the unmodified reference is only present in the build container")
def test_bench_reference_arm_uses_the_real_reference_when_present():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "48",
                        "--cpu-cores", "1"], cwd=ROOT, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-4000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
