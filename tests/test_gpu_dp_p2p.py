"""The peer-memory gradient exchange (gg_dp_allreduce, csrc/dp_allreduce.cu) needs at least two GPUs of one box: on such a box
this runs tools/dp_p2p_check.py under torchrun (bit-exact against the rank-order sum, identical on every rank, ragged ranges,
CUDA-graph replays) and tools/dp_check.py (N ranks fed the same batch follow the single-GPU loss trajectory).  Skipped on the
one-GPU box of the round-end run; the builder's multi-GPU visits are recorded under profiles/."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_bit_exact_two_ranks():
    r = _torchrun(2, "dp_p2p_check.py")
    assert r.returncode == 0 and "bit-exact" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_step_follows_single_gpu():
    r0 = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dp_check.py"), "single"], capture_output=True, text=True, timeout=600)
    assert r0.returncode == 0, r0.stderr[-2000:]
    r = _torchrun(2, "dp_check.py", "dp", port=29612)
    assert r.returncode == 0 and "worst rel loss difference" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
