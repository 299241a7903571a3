"""The learner-rank gather on hardware: NCCL gather / all-gather of a sharded batch and
the fused peer-memory epilogue, against the unsharded batch.  Needs >= 2 GPUs (gpurun --gpus 2);
the host-side logic of the same helpers runs on CPU with gloo in tests/test_dist_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_to_learner_nccl_and_peer_epilogue():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0 and "DIST_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
