"""Host-side glue of the multi-GPU path (one process per GPU, SURVEY §8e / DESIGN.md §7).

The data path is in libmdb200.so (multigpu.cu): peer stores over NVLink for the halos, ncclAllReduce for the Krylov
scalars.  What lives here is the out-of-band plumbing only: the slab arithmetic every rank evaluates identically, and
the exchange of the NCCL unique id, of the halo-window descriptors and of the owned-row counts over
``torch.distributed`` (any backend: nccl on the GPU box, gloo in the CPU tests).
"""
from types import SimpleNamespace

import numpy as np


def slab(global_n, rank, nranks, lower=None, upper=None):
    """Slab of rank `rank` when the slowest axis of a `global_n` mesh is cut into `nranks` contiguous pieces.

    Rank r owns the cell layers [c0, c1) and the vertex planes [c0, c1) (the last rank also the top plane).  Its LOCAL
    mesh additionally contains one ghost cell layer below (c0 - 1) unless it is the first rank — the convention
    mdb_mesh_partition documents — so that all cells around an owned vertex are local.
    """
    dim = len(global_n)
    lower = [0.0] * dim if lower is None else list(lower)
    upper = [1.0] * dim if upper is None else list(upper)
    pa = dim - 1
    n = int(global_n[pa])
    if not (1 <= nranks <= n):
        raise ValueError("cannot cut %d cell layers into %d slabs" % (n, nranks))
    c0 = (n * rank) // nranks
    c1 = (n * (rank + 1)) // nranks
    start = c0 - (1 if rank > 0 else 0)
    local_n = list(int(v) for v in global_n)
    local_n[pa] = c1 - start
    offset = [0] * dim
    offset[pa] = start
    h = (upper[pa] - lower[pa]) / n
    llo, lup = list(lower), list(upper)
    llo[pa] = lower[pa] + start * h
    lup[pa] = lower[pa] + c1 * h
    return SimpleNamespace(dim=dim, axis=pa, global_n=tuple(int(v) for v in global_n), local_n=tuple(local_n), cell_offset=tuple(offset),
                           owned_layers=(c0, c1), ghost_below=rank > 0, ghost_plane_above=rank < nranks - 1,
                           local_lower=tuple(llo), local_upper=tuple(lup), global_lower=tuple(lower), global_upper=tuple(upper),
                           rank=rank, nranks=nranks)


def apply_slab(api, s):
    """mesh_structured + mesh_partition for slab `s` on `api` (a Mdb context or the test oracle)."""
    api.mesh_structured(s.local_n, s.local_lower, s.local_upper)
    api.mesh_partition(s.global_n, s.cell_offset, s.global_lower, s.global_upper)


def owned_plane_range(s, k=1):
    """Lattice planes [lo, hi) along the cut axis that slab `s` owns for a Qk space (same rule as multigpu.cu)."""
    nloc = s.local_n[s.axis]
    lo = k if s.ghost_below else 0
    hi = k * nloc if s.ghost_plane_above else k * nloc + 1
    return lo, hi


def global_child_bases(per_rank_child_counts, rank):
    """Global container index of this rank's first owned DOF of every child block.

    per_rank_child_counts[r][c] = owned rows of child c on rank r.  The global numbering is lexicographic per child block
    with the cut axis slowest, so the owned rows of a child on rank r follow those of ranks < r.
    """
    counts = np.asarray(per_rank_child_counts, dtype=np.int64)
    child_total = counts.sum(axis=0)
    child_offset = np.concatenate([[0], np.cumsum(child_total)[:-1]])
    below = counts[:rank].sum(axis=0) if rank > 0 else np.zeros(counts.shape[1], dtype=np.int64)
    return child_offset + below, int(child_total.sum())


def connect(dev, dist, rank, nranks):
    """Create the NCCL communicator of `dev` and map the neighbours' halo windows.  `dist` is an initialised
    torch.distributed; call after dev.finalize()."""
    uid = [dev.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dev.comm_init(nranks, rank, uid[0])
    blob, _ = dev.halo_export()
    blobs = [None] * nranks
    dist.all_gather_object(blobs, blob)
    dev.halo_import(rank - 1, blobs[rank - 1] if rank > 0 else None, rank + 1, blobs[rank + 1] if rank < nranks - 1 else None)
    counts = [None] * nranks
    dist.all_gather_object(counts, [int(v) for v in dev.owned_counts()])
    bases, total = global_child_bases(counts, rank)
    return SimpleNamespace(child_global_base=bases, global_ndof=total, counts=counts)
