"""Hardware multi-rank correctness: NCCL ranks on real GPUs.  Skipped on a one-GPU box
(the CPU suite covers the host logic with gloo, tests/test_dist_cpu.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_nccl_ranks_equal_one_rank_bit_for_bit():
    """Sharded render (even / ragged / multi-view split) + all-gather == the frame rendered by one rank; the
    data-parallel training step (all-reduce overlapped with the fine backward, and blocking) leaves identical
    parameters on every rank, equal to the single-rank computation of the summed gradient; per-rank batches differ."""
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("RESULT ")]
    assert lines, (proc.stdout[-1500:], proc.stderr[-3000:])
    r = json.loads(lines[-1][7:])
    print(r)
    assert r["world"] == world
    bad = [k for k, v in r.items() if isinstance(v, bool) and not v]
    assert not bad, (bad, r)
