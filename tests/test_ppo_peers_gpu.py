"""The PPO gradient exchange over peer memory (dyros_ppo_reduce_peers) against NCCL's all-reduce:
needs two GPUs on the node (skipped otherwise; the driver's one-GPU run skips it, `gpurun --gpus 2` runs it)."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_exchange_equals_nccl_allreduce():
    """Two ranks, different synthetic rollouts per rank, 20 optimiser steps each way (eager NCCL, CUDA-graphed peer path):
    parameters identical across the ranks in both modes, and equal between the modes to float rounding (the sums are
    formed in rank order here, in NCCL's order there)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "ppo_peer_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=420)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-3000:]
    assert len(re.findall(r"nccl steps 20 identical across ranks True", text)) == 2, text[-3000:]
    assert len(re.findall(r"peer steps 20 identical across ranks True", text)) == 2, text[-3000:]
    assert len(re.findall(r"peer2 steps 20 identical across ranks True", text)) == 2, text[-3000:]   # reduce-scatter + all-gather
    rels = [float(x) for x in re.findall(r"peer2? vs nccl max \|diff\| \S+ rel (\S+)", text)]
    assert len(rels) == 4 and max(rels) < 1e-6, text[-3000:]
