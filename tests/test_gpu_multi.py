"""Two GPUs on one box: the column-sharded solvers (NCCL all-reduce / reduce-scatter / all-gather inside the library)
match the single-process CPU oracle. Skipped on a single-GPU box; run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_rank_nmf_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(HERE, "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_RESULT OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
