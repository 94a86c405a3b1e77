"""Real multi-GPU check (needs >= 2 GPUs on the box; skipped otherwise): the spatially sharded map over NCCL ranks equals the
unsharded map (tests/multi_gpu_shard_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
def test_spatially_sharded_map_over_nccl_ranks_equals_the_unsharded_map(tmp_path):
    n = min(4, _gpu_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "multi_gpu_shard_check.py"), str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    rep = open(os.path.join(tmp_path, "shard_check.txt")).read()
    assert "FAIL" not in rep and rep.count("OK") == 2, rep
