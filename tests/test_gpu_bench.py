"""The driver's own bench command lines must run to completion and must not accumulate HBM over the steps (round 1:
bench.py kept every step's data object and ran out of memory at --steps 20; PeerHalo leaked its CUDA-IPC buffers)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(cmd, timeout=900):
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stderr[-4000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def _check_line(r, steps, warmup, world):
    assert r["steps"] == steps and r["warmup"] == warmup and r["n_gpus"] == world
    for key in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "gpu_launches", "clocks", "config"):
        assert key in r, key
    assert r["gpu_launches"] > 0 and r["value"] > 0 and r["e2e"]["value"] > 0
    assert r["e2e"]["h2d_bytes_per_step"] > 0 and r["e2e"]["d2h_bytes_per_step"] > 0
    rf = r["roofline"]
    assert rf["bound"] == "hbm" and 0 < rf["frac"] < 1.2 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    m = r["hbm_allocated_after_step_gb"]
    # flat: what is allocated after the last timed step is what was allocated after the first one
    assert abs(m["last"] - m["first"]) <= 0.01 + 0.02 * m["first"], m
    assert m["max"] <= m["first"] * 1.05 + 0.01, m


def test_bench_driver_command_c1_memory_flat():
    r = _run([sys.executable, "bench.py", "--gpus", "1", "--steps", "20", "--warmup", "5", "--workload", "c1", "--no-cpu-baseline"])
    _check_line(r, 20, 5, 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_driver_command_world2_memory_flat():
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", "29517", "bench.py", "--gpus", "2", "--steps", "6", "--warmup", "2", "--workload", "c2"])
    _check_line(r, 6, 2, 2)
