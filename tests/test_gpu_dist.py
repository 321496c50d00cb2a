"""GPU, 2/4/8 ranks (skipped when the box has fewer GPUs): row-slab sharding with the peer-memory halo exchange gives
the same iterates as the single-GPU (replicated) path and as the oracle.  With 4 and 8 ranks the interior ranks have
both neighbours; the 1024-wide shape runs the 512-thread kernels of the 8-GPU 1024^3 benchmark line on several slab
levels.  (3-D shapes need shape[0] == shape[2]: the reference's operators use NX = shape[0] as the row stride,
openmg/operators.py:244-256, its restriction the C-order strides, openmg/operators.py:63-68.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


CASES = [((64, 64, 64), 3, 4096), ((128, 128, 128), 4, 1 << 16), ((256, 256), 4, 1024), ((1024, 1024), 5, 4096),
         ((1 << 18,), 10, 1024), ((512, 16, 512), 4, 1 << 14), ((1024, 16, 1024), 4, 1 << 16)]


@pytest.mark.parametrize("nproc", [2, 4, 8])
@pytest.mark.parametrize("shape,gl,agg", CASES)
def test_slab_sharding_matches_single_gpu_and_oracle(shape, gl, agg, nproc):
    if _ngpus() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    if shape[0] // (1 << (gl - 1)) < 1 or shape[0] % (2 * nproc) != 0:
        pytest.skip("leading extent does not split into %d even slabs" % nproc)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(29541 + nproc), os.path.join(ROOT, "tools", "dist_check.py"),
           "--shape"] + [str(s) for s in shape] + ["--gl", str(gl), "--agg", str(agg)] + (
               ["--oracle"] if int(np.prod(shape)) <= (1 << 22) else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
