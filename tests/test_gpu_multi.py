"""Several GPUs on one box: the column-sharded solvers (exchanges by the library's own NVLink peer-memory kernels and, for
comparison, by NCCL) match the single-process CPU oracle. Uses every GPU of the box up to 8; skipped on a single-GPU box
(run with `gpurun --gpus 2` ... `--gpus 8`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_multi_rank_nmf_matches_oracle():
    import torch
    ngpu = min(torch.cuda.device_count(), 8)
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ngpu), "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(HERE, "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_RESULT OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
