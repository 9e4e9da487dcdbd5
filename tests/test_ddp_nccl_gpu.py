"""N NCCL ranks == 1 rank on the concatenated batch (SURVEY section 4 item iv; reference DDP wrap trainer.py:122-129,
2257-2260), eager and CUDA-graph replayed.  Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests -m gpu`);
skipped on single-GPU boxes, where tests/test_ddp_cpu.py (gloo) covers the host-side logic."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_equal_one_rank_on_the_concatenated_batch():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "ddp_nccl_worker.py")], cwd=ROOT, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ddp nccl parity ok") == 2
