"""Data-parallel correctness on real GPUs (NCCL): needs >= 2 visible GPUs, skipped otherwise (`gpurun --gpus 2 -- pytest
tests/test_ddp_multigpu.py -m gpu`). The CPU side of the same logic (range / schedule tiling, gloo world 2) is in
tests/test_host_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_trainer_data_parallel_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "ddp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "DDP_WORKER_OK" in r.stdout, r.stderr[-4000:]
