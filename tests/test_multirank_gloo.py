"""World-size-2 (and 3) tests of the multi-rank host logic on CPU (`gloo`), SURVEY §8e / DESIGN.md §7.

The device data path (peer stores + NCCL) needs GPUs; what can and must be checked without them is the partition
design itself: slab arithmetic, "one ghost cell layer below" => owned rows are assembled completely without
communication, which planes are exchanged, and the composition of global indices from per-rank owned counts.  Every rank
drives the CPU oracle on its slab exactly as bench.py drives the CUDA backend, the halo exchange and the dot-product
all-reduce go through torch.distributed, and the results are compared with the single-rank oracle on the global mesh.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pkg = importlib.import_module("dune-multidomain_b200")
parallel = pkg.parallel


def test_slab_arithmetic_covers_the_mesh_once():
    for n, R in [(8, 2), (9, 2), (16, 3), (14, 7), (256, 8)]:
        owned_cells, owned_planes = [], []
        for r in range(R):
            s = parallel.slab((4, 4, n), r, R)
            c0, c1 = s.owned_layers
            owned_cells += list(range(c0, c1))
            assert s.cell_offset[2] == c0 - (1 if r > 0 else 0) and s.local_n[2] == c1 - s.cell_offset[2]
            lo, hi = parallel.owned_plane_range(s)
            owned_planes += [s.cell_offset[2] + p for p in range(lo, hi)]
            assert abs(s.local_lower[2] - s.cell_offset[2] / n) < 1e-15 and abs(s.local_upper[2] - c1 / n) < 1e-15
        assert owned_cells == list(range(n))
        assert owned_planes == list(range(n + 1))
    with pytest.raises(ValueError):
        parallel.slab((4, 4, 2), 0, 3)
    with pytest.raises(ValueError):             # the first slab must own two layers: mdb_mesh_partition tells the first rank by offset 0
        parallel.slab((4, 4, 7), 1, 7)


def test_global_child_bases():
    counts = [[10, 7], [12, 7], [11, 9]]        # [rank][child]
    for r, expect in enumerate([[0, 33], [10, 40], [22, 47]]):
        bases, total = parallel.global_child_bases(counts, r)
        assert list(bases) == expect and total == 56


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dof_planes(api, P, s):
    """For every local DOF: (child, in-plane lattice key, plane index along the cut axis).  Q1: DOFs sit on vertices."""
    ptr, dofs = api.get_dofmap()
    off = api.get_child_offsets()
    n = s.local_n
    child = np.searchsorted(off, np.arange(P.ndof), side="right") - 1
    key = np.full(P.ndof, -1, dtype=np.int64)
    plane = np.full(P.ndof, -1, dtype=np.int64)
    nx, ny = n[0] + 1, n[1] + 1
    for c in range(len(ptr) - 1):
        i, j, k = c % n[0], (c // n[0]) % n[1], c // (n[0] * n[1])
        loc = dofs[ptr[c]:ptr[c + 1]]
        assert len(loc) == 8               # one Q1 child lives on every cell of this problem
        for li, g in enumerate(loc):
            ax, ay, az = li & 1, (li >> 1) & 1, (li >> 2) & 1
            key[g] = (i + ax) + nx * (j + ay)
            plane[g] = k + az
    assert (plane >= 0).all()
    return child, key, plane


def _worker(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import orc
    import problems
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        # ---- this rank's slab, set up exactly like bench.py does for the device --------------------------------------
        s = parallel.slab(n, rank, world)
        o = orc.Oracle()
        P = problems.testpoisson(o, n, slab=s)
        child, key, plane = _dof_planes(o, P, s)
        lo, hi = parallel.owned_plane_range(s)
        owned = (plane >= lo) & (plane < hi)
        nchild = o.nchildren
        counts = [None] * world
        dist.all_gather_object(counts, [int((owned & (child == c)).sum()) for c in range(nchild)])
        bases, nglobal = parallel.global_child_bases(counts, rank)
        gidx = np.full(P.ndof, -1, dtype=np.int64)
        for c in range(nchild):
            sel = np.where(owned & (child == c))[0]           # ascending local index = ascending global index
            gidx[sel] = bases[c] + np.arange(len(sel))

        # ---- halo exchange through gloo: owners send their boundary planes, keyed by (child, in-plane position) --------
        nloc = s.local_n[2]

        def plane_dofs(p):
            sel = np.where(plane == p)[0]
            return sel[np.lexsort((key[sel], child[sel]))]

        def exchange(v):
            reqs, recvs = [], []
            if rank > 0:                                     # my first owned plane -> ghost plane above of rank-1
                t = torch.from_numpy(v[plane_dofs(1)].copy())
                reqs.append(dist.isend(t, rank - 1))
                buf = torch.empty(len(plane_dofs(0)), dtype=torch.float64)
                reqs.append(dist.irecv(buf, rank - 1)); recvs.append((plane_dofs(0), buf))
            if rank < world - 1:                             # my last owned plane -> ghost plane below of rank+1
                t2 = torch.from_numpy(v[plane_dofs(nloc - 1)].copy())
                reqs.append(dist.isend(t2, rank + 1))
                buf2 = torch.empty(len(plane_dofs(nloc)), dtype=torch.float64)
                reqs.append(dist.irecv(buf2, rank + 1)); recvs.append((plane_dofs(nloc), buf2))
            for q in reqs:
                q.wait()
            for idx, b in recvs:
                v[idx] = b.numpy()

        def dot(a, b):
            t = torch.tensor([float(np.dot(a[owned], b[owned]))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        g = gidx.astype(np.float64)
        exchange(g)                                           # ghosts learn their global index from their owner
        gfull = g.astype(np.int64)
        assert (gfull >= 0).all(), "a ghost DOF was not covered by the halo exchange"

        # ---- the global problem on one rank (reference for everything below) ------------------------------------------
        og = orc.Oracle()
        Pg = problems.testpoisson(og, n)
        assert nglobal == Pg.ndof
        ug = np.zeros(Pg.ndof); og.interpolate(Pg.sps, Pg.gfn, ug)
        og.jacobian(Pg.go, Pg.mat, ug, overwrite=True)
        rg = np.zeros(Pg.ndof); og.residual(Pg.go, ug, rg)
        Ag = og.csr(Pg.mat)

        # ---- local assembly: owned rows must be complete without any communication -------------------------------------
        u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
        assert np.array_equal(u, ug[gfull]), "interpolation differs from the global one"
        assert np.array_equal(o.get_constraints()[owned], og.get_constraints()[gfull[owned]])
        o.jacobian(P.go, P.mat, u, overwrite=True)
        r = np.zeros(P.ndof); o.residual(P.go, u, r)
        A = o.csr(P.mat)
        scale = np.abs(Ag.data).max()
        for row in np.where(owned)[0]:
            a, b = A.indptr[row], A.indptr[row + 1]
            ga, gb = Ag.indptr[gfull[row]], Ag.indptr[gfull[row] + 1]
            cols = gfull[A.indices[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb]), "pattern of an owned row differs from the global matrix"
            assert np.abs(A.data[a:b][order] - Ag.data[ga:gb]).max() <= 1e-13 * scale
        assert np.abs(r[owned] - rg[gfull[owned]]).max() <= 1e-13 * np.abs(rg).max()

        # ---- distributed Jacobi-CG (the iteration of solver.cu) vs the global direct solve ------------------------------
        dinv = 1.0 / A.diagonal()
        x = np.zeros(P.ndof); res = r.copy(); p = np.zeros(P.ndof)
        rho = dot(res, dinv * res); rr0 = dot(res, res); beta = 0.0
        for it in range(500):
            p = beta * p + dinv * res
            exchange(p)
            q = A @ p
            lam = rho / dot(p, q)
            x += lam * p
            res -= lam * q
            rho_new = dot(res, dinv * res)
            beta, rho = rho_new / rho, rho_new
            if dot(res, res) <= 1e-26 * rr0:
                break
        zg = spla.spsolve(sp.csc_matrix(Ag), rg)
        err = np.abs(x[owned] - zg[gfull[owned]]).max() / np.abs(zg).max()
        assert err < 1e-10, err
        # ---- the same iteration with block-Jacobi SeqSSOR(1) (precond.cu on a partition): each rank sweeps its
        # owned x owned block only; the preconditioner changes with the rank count, the answer does not ------------------------
        ow = np.where(owned)[0]
        B = A[ow][:, ow].tocsr(); B.sort_indices()
        bd = B.diagonal()

        def block_ssor(d):
            v = np.zeros(P.ndof); w = np.zeros(len(ow)); dd = d[ow]
            for order in (range(len(ow)), range(len(ow) - 1, -1, -1)):
                for i in order:
                    a, b = B.indptr[i], B.indptr[i + 1]
                    w[i] += (dd[i] - B.data[a:b] @ w[B.indices[a:b]]) / bd[i]
            v[ow] = w
            return v
        x = np.zeros(P.ndof); res = r.copy(); p = np.zeros(P.ndof)
        zv = block_ssor(res); rho = dot(res, zv); beta = 0.0
        for it2 in range(200):
            p = beta * p + zv
            exchange(p)
            q = A @ p
            lam = rho / dot(p, q)
            x += lam * p
            res -= lam * q
            zv = block_ssor(res)
            rho_new = dot(res, zv)
            beta, rho = rho_new / rho, rho_new
            if dot(res, res) <= 1e-26 * rr0:
                break
        err2 = np.abs(x[owned] - zg[gfull[owned]]).max() / np.abs(zg).max()
        assert err2 < 1e-10 and it2 + 1 < it + 1, (err2, it2, it)      # converges to the global solution, in fewer steps than Jacobi
        ret[rank] = (int(owned.sum()), it + 1, float(err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, (4, 6, 8)), (3, (4, 4, 9))])
def test_partitioned_assembly_and_solve_match_global(world, n):
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n, ret), nprocs=world, join=True)
    assert len(ret) == world
    total = sum(v[0] for v in ret.values())
    assert total == (n[0] + 1) * (n[2] + 1) * (n[1] + 2)     # (N+2)(N+1)^2-type count: interface vertices belong to both blocks
    assert len({v[1] for v in ret.values()}) == 1             # every rank took the same number of iterations


# ---- unstructured meshes: recursive coordinate bisection, rows owned by the lowest rank touching the vertex ----------
def test_um_partition_invariants():
    import um_problems as ump
    xy, conn = ump.square_mesh(9, 7, seed=4)
    for R in (1, 2, 3, 5):
        cell_rank, owner = parallel.um_partition(xy, conn, R)
        counts = np.bincount(cell_rank, minlength=R)
        assert counts.sum() == conn.shape[0] and counts.max() - counts.min() <= R      # balanced up to the rounding of each cut
        assert owner.min() >= 0 and owner.max() < R
        for v in range(xy.shape[0]):
            assert owner[v] == cell_rank[(conn == v).any(axis=1)].min()                 # lowest rank among the vertex's cells
        covered = np.zeros(xy.shape[0], dtype=int)
        for r in range(R):
            cells, verts, lconn = parallel.um_local_mesh(conn, owner, r)
            assert np.array_equal(verts[lconn], conn[cells]) and np.all(np.diff(cells) > 0) and np.all(np.diff(verts) > 0)
            mine = np.nonzero(owner == r)[0]
            covered[mine] += 1
            for v in mine:                                                              # the full cell ring of every owned vertex is local
                assert np.isin(np.nonzero((conn == v).any(axis=1))[0], cells).all()
        assert np.all(covered == 1)


def _um_worker(rank, world, port, ret):
    import torch.distributed as dist
    import torch
    import um_problems as ump
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        capi = ump.capi
        xy, conn = ump.square_mesh(8, 6, seed=17)
        sd = ump.halves(xy, conn)
        kw = dict(cvcf=True, k=2.0)
        # the global problem (every rank computes it: the reference for its own rows)
        og = ump.oracle(); Hg = ump.setup(og, xy, conn, sd, **kw)
        og.jacobian(Hg["go"], Hg["mat"], Hg["x0"], overwrite=True)
        rg = og.residual(Hg["go"], Hg["x0"], np.zeros(Hg["ndof"]))
        rpg, cig = og.matrix_get_pattern(Hg["mat"]); vg = og.matrix_get_values(Hg["mat"])
        offg = og.get_child_offsets()
        # the local problem on the rank's part + ghost cells
        cell_rank, owner = parallel.um_partition(xy, conn, world)
        cells, verts, lconn = parallel.um_local_mesh(conn, owner, rank)
        ol = ump.oracle(); Hl = ump.setup(ol, xy[verts], lconn, sd[cells], **kw)
        offl = ol.get_child_offsets()
        # local DOF -> global DOF through (child, vertex); owned = the vertex is owned by this rank
        l2g = -np.ones(Hl["ndof"], dtype=np.int64); owned = np.zeros(Hl["ndof"], dtype=bool)
        for k in range(2):
            vl = ol.children[k]["v2d"]; vgl = og.children[k]["v2d"]
            for lv in np.nonzero(vl >= 0)[0]:
                l2g[offl[k] + vl[lv]] = offg[k] + vgl[verts[lv]]
                owned[offl[k] + vl[lv]] = owner[verts[lv]] == rank
        assert (l2g >= 0).all()
        # the linearisation point: the global one restricted (ghost values come from their owners in a real run)
        xl = Hg["x0"][l2g]
        ol.jacobian(Hl["go"], Hl["mat"], xl, overwrite=True)
        rl = ol.residual(Hl["go"], xl, np.zeros(Hl["ndof"]))
        rpl, cil = ol.matrix_get_pattern(Hl["mat"]); vl_ = ol.matrix_get_values(Hl["mat"])
        scale = np.abs(vg).max()
        conl, cong = ol.get_constraints(), og.get_constraints()
        for i in np.nonzero(owned)[0]:
            g = l2g[i]
            cols = l2g[cil[rpl[i]:rpl[i + 1]]]
            assert np.array_equal(cols, cig[rpg[g]:rpg[g + 1]]), "pattern of an owned row differs from the global one"   # monotone maps keep the order
            assert np.abs(vl_[rpl[i]:rpl[i + 1]] - vg[rpg[g]:rpg[g + 1]]).max() <= 1e-13 * scale
            assert abs(rl[i] - rg[g]) <= 1e-13 * np.abs(rg).max() and conl[i] == cong[g]
        # every global row is owned exactly once
        t = torch.zeros(Hg["ndof"], dtype=torch.float64); t[l2g[owned]] = 1.0
        dist.all_reduce(t)
        assert bool((t == 1.0).all())
        ret[rank] = int(owned.sum())
        # ---- distributed Jacobi-CG on the owned rows: ghost entries come from their owners, dots are all-reduced -------------------
        import scipy.sparse as sps
        import scipy.sparse.linalg as spla
        Al = sps.csr_matrix((vl_, cil, rpl), shape=(Hl["ndof"],) * 2)
        Ag = sps.csr_matrix((vg, cig, rpg), shape=(Hg["ndof"],) * 2)
        zg = spla.spsolve(Ag.tocsc(), rg)

        def exchange(v):                                        # test plumbing: owners publish, everybody reads its ghosts
            t = torch.zeros(Hg["ndof"], dtype=torch.float64)
            t[l2g[owned]] = torch.from_numpy(v[owned].copy())
            dist.all_reduce(t)
            return t.numpy()[l2g]

        def dot(a, b):
            t = torch.tensor([float(np.dot(a[owned], b[owned]))], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t.item())

        dinv = np.where(owned, 1.0 / np.where(Al.diagonal() != 0.0, Al.diagonal(), 1.0), 0.0)
        b = np.where(owned, rl, 0.0)
        z = np.zeros(Hl["ndof"]); r = b.copy(); p = np.zeros_like(z)
        rho_old, it, d0 = 1.0, 0, np.sqrt(dot(b, b))
        while it < 2000:
            y = dinv * r
            rho = dot(r, y)
            p = y + (rho / rho_old) * p if it > 0 else y.copy()
            q = np.where(owned, Al @ exchange(p), 0.0)
            alpha = rho / dot(p, q)
            z += alpha * p; r -= alpha * q
            rho_old = rho; it += 1
            if np.sqrt(dot(r, r)) <= 1e-13 * d0:
                break
        assert it < 2000, "distributed CG did not converge"
        assert np.abs(z[owned] - zg[l2g[owned]]).max() <= 1e-9 * np.abs(zg).max()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_um_partitioned_assembly_owned_rows_match_global(world):
    """Unstructured partition design on CPU ranks: each rank assembles its part plus ghost cells with the oracle; the rows it owns —
    pattern, FD-coupling Jacobian entries, residual, constraint flags — equal the global ones without any communication."""
    import torch.multiprocessing as mp
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_um_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world and sum(ret.values()) > 0


# ---- mdb_partition_lists plumbing (parallel.um_rank_problem / um_connect) without a GPU ------------------------------------------
class _RecordingDev:
    """Stands in for capi.Mdb in um_connect: records the lists instead of uploading them."""

    def __init__(self, rank):
        self.rank = rank

    def partition_lists(self, owned, neigh, send_ptr, send_idx, recv_ptr, recv_idx):
        self.owned, self.neigh = np.asarray(owned), list(neigh)
        self.send_ptr, self.send_idx, self.recv_ptr, self.recv_idx = list(send_ptr), np.asarray(send_idx), list(recv_ptr), np.asarray(recv_idx)

    def nccl_unique_id(self):
        return b"id"

    def comm_init(self, nranks, rank, uid):
        assert uid == b"id"

    def halo_export(self):
        return b"window-of-%d" % self.rank, 0

    def halo_import_lists(self, blobs, caps, offsets, slots):
        self.blobs, self.caps, self.offsets, self.slots = list(blobs), list(caps), list(offsets), list(slots)


def _um_lists_worker(rank, world, port, ret):
    import torch.distributed as dist
    import um_problems as ump
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        xy, conn = ump.square_mesh(14, 11, seed=5)
        sd = ump.halves(xy, conn)
        R = parallel.um_rank_problem(xy, conn, sd, world, rank, [0, 1])
        dev = _RecordingDev(rank)
        info = parallel.um_connect(dev, dist, rank, world, R.global_dof, R.owner)
        # emulate one exchange of the vector x[g] = f(global index): what the neighbours write into my window, at the offsets they were told
        f = lambda g: np.sin(0.37 * g) + 1e-3 * g
        x = np.where(R.owned, f(R.global_dof), np.nan)
        packets = [None] * world
        dist.all_gather_object(packets, {q: (dev.offsets[i], [float(v) for v in x[dev.send_idx[dev.send_ptr[i]:dev.send_ptr[i + 1]]]]) for i, q in enumerate(dev.neigh)})
        window = np.full(max(dev.recv_ptr[-1], 1), np.nan)
        for q in range(world):
            if q != rank and rank in packets[q]:
                off, vals = packets[q][rank]
                window[off:off + len(vals)] = vals
        x[dev.recv_idx] = window[:dev.recv_ptr[-1]]
        ok = bool(np.allclose(x, f(R.global_dof), rtol=0, atol=0))
        # every global DOF is owned exactly once
        owned_ids = [None] * world
        dist.all_gather_object(owned_ids, [int(g) for g in R.global_dof[R.owned]])
        allg = np.sort(np.concatenate([np.asarray(o, dtype=np.int64) for o in owned_ids]))
        once = bool(np.array_equal(allg, np.arange(R.global_ndof)))
        sym = all(dev.blobs[i] == b"window-of-%d" % q for i, q in enumerate(dev.neigh))
        ret[rank] = (ok, once, sym, info.send_count, info.recv_count)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_um_halo_lists_exchange_every_ghost_from_its_owner(world):
    import torch.multiprocessing as mp
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_um_lists_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, once, sym, ns, nr = ret[r]
        assert ok and once and sym and ns > 0 and nr > 0
    assert sum(ret[r][3] for r in range(world)) == sum(ret[r][4] for r in range(world))      # everything sent is received


@pytest.mark.parametrize("n,k,world", [((6, 9), 2, 3), ((3, 4, 7), 1, 2)])
def test_cell_blocked_slab_ids(n, k, world):
    """parallel.slab2 + cell_blocked_ids (QkDG spaces on a partition, index-list halos): with the oracle as the context of every rank,
    the ids agree with the one-rank numbering through the cell maps, every global DOF is owned exactly once, every ghost DOF is owned by
    a neighbour, and the owned rows of the local matrices equal the global rows (no matrix communication)."""
    import orc
    import problems
    og = orc.Oracle()
    Pg = problems.dirichlet_neumann_dg(og, n, k=k)
    x = np.random.default_rng(1).uniform(-1, 1, Pg.ndof)
    og.jacobian(Pg.go, Pg.mat, x, overwrite=True)
    Ag = og.csr(Pg.mat)
    pg, dg = og.get_dofmap()
    seen = np.zeros(Pg.ndof, dtype=int)
    idmap = {}
    for rank in range(world):
        s = parallel.slab2(n, rank, world)
        ol = orc.Oracle()
        P = problems.dirichlet_neumann_dg(ol, n, k=k, box=s)
        pl, dl = ol.get_dofmap()
        ids, owner = parallel.cell_blocked_ids(pl, dl, P.ndof, s)
        gidx = parallel.slab_global_dofs(pl, dl, P.ndof, s, pg, dg)
        owned = owner == rank
        seen[gidx[owned]] += 1
        assert set(np.unique(owner[~owned])) <= {rank - 1, rank + 1}
        for i, g in zip(ids, gidx):                     # a rank-independent id: the same global DOF carries the same id on every rank
            assert idmap.setdefault(int(i), int(g)) == int(g)
        ol.jacobian(P.go, P.mat, np.ascontiguousarray(x[gidx]), overwrite=True)
        Al = ol.csr(P.mat)
        for row in np.where(owned)[0]:
            a, b = Al.indptr[row], Al.indptr[row + 1]
            ga, gb = Ag.indptr[gidx[row]], Ag.indptr[gidx[row] + 1]
            cols = gidx[Al.indices[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb])
            assert np.abs(Al.data[a:b][order] - Ag.data[ga:gb]).max() <= 1e-12 * np.abs(Ag.data).max()
        ol.close()
    assert (seen == 1).all()
    og.close()
