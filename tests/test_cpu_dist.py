"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path — the slab partition
(omg_partition, pure host code in the C-ABI library) and the gather helpers."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from openmg_b200 import dist as odist


def hierarchy_rows(shape, nlev):
    lead = [shape[0] >> l for l in range(nlev)]
    rows = [int(np.prod([s >> l for s in shape])) for l in range(nlev)]
    return lead, rows


def test_partition_single_rank_is_identity():
    lead, rows = hierarchy_rows((64, 64, 64), 4)
    ld, row0, nloc = odist.partition(lead, rows, [1, 1, 1, 0], 1, 0)
    assert ld == 0 and list(row0) == [0] * 4 and list(nloc) == rows


def test_partition_covers_rows_and_keeps_aggregates_whole():
    shape, nlev = (512, 512, 512), 6
    lead, rows = hierarchy_rows(shape, nlev)
    for nranks in (2, 4, 8):
        parts = [odist.partition(lead, rows, [1, 1, 1, 1, 1, 0], nranks, r) for r in range(nranks)]
        ld = parts[0][0]
        assert ld == 3                                  # 512^3, 256^3, 128^3 sharded; 64^3 and below replicated
        for l in range(nlev):
            r0 = [int(p[1][l]) for p in parts]
            nl = [int(p[2][l]) for p in parts]
            if l < ld:
                assert r0[0] == 0 and all(r0[i + 1] == r0[i] + nl[i] for i in range(nranks - 1))
                assert r0[-1] + nl[-1] == rows[l]
                plane = rows[l] // lead[l]
                assert all(v % (2 * plane) == 0 for v in r0)          # cuts on even planes: no aggregate straddles
                if l + 1 < ld:                                        # coarse slab = image of the fine slab
                    assert [v // 8 for v in r0] == [int(p[1][l + 1]) for p in parts]
            else:
                assert r0 == [0] * nranks and nl == [rows[l]] * nranks
    # threshold: nothing below it is sharded
    ld, _, _ = odist.partition(lead, rows, [1, 1, 1, 1, 1, 0], 8, 3, agglomerate_below=1 << 25)
    assert ld == 1
    # leading extent must divide evenly at the transition level
    lead3, rows3 = hierarchy_rows((48, 48, 48), 3)
    ld, _, _ = odist.partition(lead3, rows3, [1, 1, 0], 8, 0, agglomerate_below=16)
    assert ld == 1 and lead3[1] % 8 == 0


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from openmg_b200 import dist as odist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
shape, nlev = (16, 16, 16), 3
lead = [shape[0] >> l for l in range(nlev)]
rows = [int(np.prod([s >> l for s in shape])) for l in range(nlev)]
ld, row0, nloc = odist.partition(lead, rows, [1, 1, 0], world, rank, agglomerate_below=64)
assert ld == 2, ld
full = np.random.RandomState(0).random_sample(rows[0])
mine = odist.local_slice(full, int(row0[0]), int(nloc[0]))
meta = [None] * world
dist.all_gather_object(meta, (int(row0[0]), int(nloc[0])))
got = odist.allgather_rows(dist, mine * 2.0, [m[0] for m in meta], [m[1] for m in meta])
assert np.array_equal(got, full * 2.0)
# halo bookkeeping: the neighbour's boundary plane is what my halo must receive
plane = rows[0] // lead[0]
lo_needed = full[int(row0[0]) - plane:int(row0[0])] if rank > 0 else None
pieces = [None] * world
dist.all_gather_object(pieces, mine[-plane:])
if rank > 0:
    assert np.array_equal(pieces[rank - 1], lo_needed)
dist.barrier()
dist.destroy_process_group()
print("rank %%d ok" %% rank)
'''


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


def test_partition_properties_across_dimensions_and_rank_counts():
    """1-D / 2-D / 3-D hierarchies on 2..8 ranks: slab levels tile the rows exactly, cuts fall on even leading
    indices (no aggregate straddles two ranks), and the coarse slab of a rank is the image of its fine slab."""
    cases = [((1 << 16,), 10), ((4096, 4096), 6), ((512, 1024, 512), 5), ((96, 96, 96), 3), ((1024, 1024, 1024), 6)]
    for shape, nlev in cases:
        lead, rows = hierarchy_rows(shape, nlev)
        k = 2 ** len(shape)
        for nranks in (2, 3, 4, 6, 8):
            regular = [1] * (nlev - 1) + [0]
            parts = [odist.partition(lead, rows, regular, nranks, r, agglomerate_below=1 << 12) for r in range(nranks)]
            ld = parts[0][0]
            assert all(p[0] == ld for p in parts)
            for l in range(nlev):
                r0 = [int(p[1][l]) for p in parts]
                nl = [int(p[2][l]) for p in parts]
                if l >= ld:
                    assert r0 == [0] * nranks and nl == [rows[l]] * nranks
                    continue
                assert rows[l] > (1 << 12)
                assert r0[0] == 0 and r0[-1] + nl[-1] == rows[l] and all(n > 0 for n in nl)
                assert all(r0[i + 1] == r0[i] + nl[i] for i in range(nranks - 1))
                stride = rows[l] // lead[l]                       # rows per leading index
                assert all(v % (2 * stride) == 0 for v in r0 + nl)
                if l + 1 < ld:
                    assert [v // k for v in r0] == [int(p[1][l + 1]) for p in parts]
                    assert [v // k for v in nl] == [int(p[2][l + 1]) for p in parts]
