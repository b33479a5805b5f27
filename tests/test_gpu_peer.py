"""Multi-GPU exchanges over NVLink peer memory (csrc/peer.cu) on real hardware: needs >= 2 GPUs on the box (skipped on the
single-GPU test box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_peer.py -m gpu` runs it)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_exchanges_match_nccl_and_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("PEER_RESULT ")]
    assert line, r.stdout[-2000:]
    out = json.loads(line[-1][len("PEER_RESULT "):])
    print(out)
    assert out["fail_flag"] == 0
    assert out["bit_identical"]
    assert out["allreduce_err"] < 1e-2              # sums of ~1e3-sized floats in a different association
    assert out["loss_vs_nccl"] < 1e-5 and out["grad_vs_nccl"] < 1e-4
    assert out["loss_vs_single"] < 1e-5 and out["grad_vs_single"] < 1e-4
