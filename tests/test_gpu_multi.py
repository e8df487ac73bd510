"""The multi-process GPU path on real hardware: tools/mg_check.py under torchrun, one rank per GPU over NCCL --
row-sharded sgemm with the C stripes pushed from the epilogue (CUDA-IPC peers, symmetric-memory unicast, NVLS multicast)
against plain stripes + ncclAllGather (bit-identical) and float64 products; sharded-axis reduceSum / reduceMax /
reduceArgMax merged over NCCL against the oracle on the whole matrix, NaN at local position 0 of a later rank included
(PhpMath.php:114-130).  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a single-GPU box, where tests/test_sharding_gloo.py
(CPU, gloo, world 2) covers the host logic and tests/test_gpu_allgather.py the epilogue stores.  pytest -m gpu"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_over_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "mg_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout[-6000:])
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-4000:]
    assert "FAIL" not in p.stdout
