"""Multi-GPU plumbing on the Python side: one process per GPU (torchrun), torch.distributed
only for rendezvous (broadcast of the NCCL unique id) and host-side gathers; the data path
(halo exchange, all-gather of the coarse right-hand side, norm all-reduce) is NCCL inside
libomg_b200.so."""
import ctypes

import numpy as np

from . import _lib


def partition(level_lead, level_rows, level_regular, nranks, rank, agglomerate_below=1 << 19):
    """Host-only: (first_replicated, row0[], nloc[]) of omg_partition (no GPU needed)."""
    L = _lib.load()
    n = len(level_rows)
    lead = np.ascontiguousarray(level_lead, dtype=np.int64)
    rows = np.ascontiguousarray(level_rows, dtype=np.int64)
    reg = np.ascontiguousarray(level_regular, dtype=np.int32)
    row0 = np.zeros(n, np.int64)
    nloc = np.zeros(n, np.int64)
    ld = ctypes.c_int32()
    rc = L.omg_partition(n, _lib.i64(lead), _lib.i64(rows), _lib.i32(reg), int(nranks), int(rank),
                         int(agglomerate_below), ctypes.byref(ld), _lib.i64(row0), _lib.i64(nloc))
    if rc != 0:
        raise ValueError(L.omg_last_error().decode())
    return ld.value, row0, nloc


_INITED = None


def active():
    """torch.distributed when this process is one rank of an initialised multi-rank group (torchrun), else None.
    torch is only imported if the launcher's environment says there is such a group."""
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None
    try:
        import torch.distributed as dist
    except Exception:  # noqa: BLE001
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def init_from_torch(dist):
    """Create the library's NCCL communicator for the ranks of an initialised torch.distributed group (once)."""
    global _INITED
    rank, world = dist.get_rank(), dist.get_world_size()
    if _INITED == (rank, world):
        return rank, world
    L = _lib.lib()
    box = [None]
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        _lib.check(L.omg_nccl_unique_id(buf))
        box[0] = bytes(buf.raw)
    dist.broadcast_object_list(box, src=0)
    _lib.check(L.omg_dist_init(rank, world, box[0]))
    _INITED = (rank, world)
    return rank, world


def local_slice(vec, row0, nloc):
    return np.ascontiguousarray(np.asarray(vec).ravel()[row0:row0 + nloc])


def allgather_rows(dist, local, row0s, nlocs):
    """Assemble a global vector from per-rank row slices (host side, any backend)."""
    world = dist.get_world_size()
    pieces = [None] * world
    dist.all_gather_object(pieces, np.asarray(local))
    out = np.zeros(int(sum(nlocs)), dtype=np.float64)
    for r in range(world):
        out[int(row0s[r]):int(row0s[r]) + int(nlocs[r])] = pieces[r]
    return out


def gather_solution(dist, hierarchy, x, level=0):
    """Full level vector on every rank from the rows each rank owns."""
    row0, nloc, slab = hierarchy.local_range(level)
    if not slab:
        return np.asarray(x)
    world = dist.get_world_size()
    if dist.get_backend() == "nccl":
        # slabs are equal-sized and ordered by rank (omg_partition): one device all-gather
        import torch
        mine = torch.from_numpy(local_slice(x, row0, nloc)).cuda()
        full = torch.empty(world * nloc, dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(full, mine)
        return full.cpu().numpy()
    meta = [None] * world
    dist.all_gather_object(meta, (row0, nloc))
    return allgather_rows(dist, local_slice(x, row0, nloc), [m[0] for m in meta], [m[1] for m in meta])
