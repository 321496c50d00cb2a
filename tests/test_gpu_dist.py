"""GPU, 2 ranks (skipped on a single-GPU box): row-slab sharding with NCCL halo exchange gives
the same iterates as the single-GPU path."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.parametrize("shape,gl,agg", [((64, 64, 64), 3, 4096), ((128, 128, 128), 4, 1 << 16),
                                          ((256, 256), 4, 1024), ((1024, 1024), 5, 4096), ((65536,), 8, 1024)])
def test_slab_sharding_matches_single_gpu(shape, gl, agg):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py"),
           "--shape"] + [str(s) for s in shape] + ["--gl", str(gl), "--agg", str(agg)] + (
               ["--oracle"] if int(__import__("numpy").prod(shape)) <= 262144 else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
