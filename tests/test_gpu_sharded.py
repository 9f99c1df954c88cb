"""Time-sharded path on real GPUs (needs >= 2 GPUs in the box; skipped otherwise): NCCL run under torchrun,
compared id for id with the single-GPU result (tools/check_sharded.py)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

REPO = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.parametrize("frames,eps_time", [("10", "2.0"), ("12", "5.0")])          # defaults; BASELINE config 5's 5-frame halo
def test_sharded_nccl_equals_single_gpu(frames, eps_time):
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(REPO / "tools" / "check_sharded.py"), frames, eps_time]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "IDENTICAL to single GPU" in res.stdout
