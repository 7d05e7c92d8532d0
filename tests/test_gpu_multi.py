"""GPU tests of the multi-GPU path (multigpu.cu) through the C ABI.

* one GPU: two contexts of one process act as two ranks.  Their windows are exchanged as raw pointers (the 80-byte
  descriptor carries a per-process token), pushes and waits are issued in an order in which no kernel ever waits for a kernel that
  has not been launched yet, so this is safe on a single device.  Checks: partitioned assembly vs the oracle's global
  matrix through the exported global indices, halo push / wait, distributed SpMV.
* two or more GPUs (skipped otherwise): one process per GPU, NCCL + cudaIpc windows, full Krylov solves against the
  oracle's solution of the global problem.
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
capi, parallel = pkg.capi, pkg.parallel

pytestmark = pytest.mark.gpu


def _global_oracle(n):
    import orc
    import problems
    og = orc.Oracle()
    Pg = problems.testpoisson(og, n)
    ug = np.zeros(Pg.ndof)
    og.interpolate(Pg.sps, Pg.gfn, ug)
    og.jacobian(Pg.go, Pg.mat, ug, overwrite=True)
    rg = np.zeros(Pg.ndof)
    og.residual(Pg.go, ug, rg)
    return og, Pg, ug, rg


@pytest.mark.parametrize("n", [(4, 6, 8), (5, 4, 7)])
def test_two_contexts_on_one_gpu(n):
    import torch
    import problems
    R = 2
    devs = [pkg.Mdb(0) for _ in range(R)]
    Ps = [problems.testpoisson(devs[r], n, slab=parallel.slab(n, r, R)) for r in range(R)]
    counts = [[int(v) for v in d.owned_counts()] for d in devs]
    og, Pg, ug, rg = _global_oracle(n)
    Ag = og.csr(Pg.mat)
    assert sum(sum(c) for c in counts) == Pg.ndof
    blobs = [d.halo_export()[0] for d in devs]
    devs[0].halo_import(-1, None, 1, blobs[1])
    devs[1].halo_import(0, blobs[0], -1, None)
    gfull, owned, xs = [], [], []
    for r in range(R):
        bases, total = parallel.global_child_bases(counts, r)
        assert total == Pg.ndof
        f, g = devs[r].get_ownership(bases)
        assert int(f.sum()) == devs[r].owned_rows()[0] == sum(counts[r])
        owned.append(f.astype(bool))
        xs.append(torch.from_numpy(g.astype(np.float64)).cuda())
    torch.cuda.synchronize()
    # exchange the global indices themselves: afterwards every ghost entry knows its owner's global index
    for r in range(R):
        devs[r].halo_push(xs[r])
    for r in range(R):
        devs[r].synchronize()
    for r in range(R):
        devs[r].halo_wait(xs[r])
        devs[r].synchronize()
        g = xs[r].cpu().numpy().astype(np.int64)
        assert (g >= 0).all(), "a ghost DOF was not covered by the halo exchange"
        gfull.append(g)
    for r in range(R):
        d, P = devs[r], Ps[r]
        u = np.zeros(P.ndof)
        d.interpolate(P.sps, P.gfn, u)
        assert np.allclose(u, ug[gfull[r]], rtol=1e-14, atol=0)      # device exp() vs libm exp(): last-bit differences
        assert np.array_equal(d.get_constraints()[owned[r]], og.get_constraints()[gfull[r][owned[r]]])
        u = np.ascontiguousarray(ug[gfull[r]])         # same linearisation point as the oracle: FD Jacobians amplify one ulp of x
        d.jacobian(P.go, P.mat, u, overwrite=True)
        rp, ci = d.matrix_get_pattern(P.mat)
        va = d.matrix_get_values(P.mat)
        scale = np.abs(Ag.data).max()
        for row in np.where(owned[r])[0]:
            a, b = rp[row], rp[row + 1]
            ga, gb = Ag.indptr[gfull[r][row]], Ag.indptr[gfull[r][row] + 1]
            cols = gfull[r][ci[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb])
            assert np.abs(va[a:b][order] - Ag.data[ga:gb]).max() <= 1e-12 * scale
        res = np.zeros(P.ndof)
        d.residual(P.go, u, res)
        assert np.abs(res[owned[r]] - rg[gfull[r][owned[r]]]).max() <= 1e-12 * np.abs(rg).max()
    # halo exchange of a vector known only on owned rows (ghosts poisoned): afterwards the ghosts hold the owners' values.
    # (mdb_spmv / mdb_solve issue push + wait back to back, so they need every rank to be driven concurrently: that is
    # the multi-process test below.)
    rng = np.random.default_rng(0)
    xg = rng.uniform(-1, 1, Pg.ndof)
    xd = [torch.from_numpy(np.where(owned[r], xg[gfull[r]], 1e300)).cuda() for r in range(R)]
    torch.cuda.synchronize()
    for r in range(R):
        devs[r].halo_push(xd[r])
        devs[r].synchronize()
    for r in range(R):
        devs[r].halo_wait(xd[r])
        devs[r].synchronize()
        assert np.array_equal(xd[r].cpu().numpy(), xg[gfull[r]])
    for d in devs:
        d.close()


QK_CASES = [((capi.FE_Q1, capi.FE_Q2), (4, 5, 6)), ((capi.FE_Q2, capi.FE_Q2), (3, 4, 6)), ((capi.FE_Q2, capi.FE_Q1), (4, 3, 7))]


def _global_oracle_fe(n, fe):
    import orc
    import problems
    og = orc.Oracle()
    Pg = problems.testpoisson(og, n, fe=fe)
    ug = np.zeros(Pg.ndof)
    og.interpolate(Pg.sps, Pg.gfn, ug)
    og.jacobian(Pg.go, Pg.mat, ug, overwrite=True)
    rg = np.zeros(Pg.ndof)
    og.residual(Pg.go, ug, rg)
    return og, Pg, ug, rg


@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("fe,n", QK_CASES)
def test_two_contexts_on_one_gpu_qk(fe, n, R):
    """Slab partition with Qk (k = 2) child spaces (the shipped testpoisson is Q1 | Q2): two contexts of one process act as two ranks.
    Ownership by lattice planes (k ghost planes below, one above), halo exchange of k planes up / one plane down, and the owned rows
    of the partitioned assembly against the oracle's global matrix — the local -> global map goes through the cell DOF maps."""
    import torch
    import problems
    devs = [pkg.Mdb(0) for _ in range(R)]                 # R = 3: the middle rank has ghosts on both sides and two neighbours
    slabs = [parallel.slab(n, r, R) for r in range(R)]
    Ps = [problems.testpoisson(devs[r], n, fe=fe, slab=slabs[r]) for r in range(R)]
    og, Pg, ug, rg = _global_oracle_fe(n, fe)
    Ag = og.csr(Pg.mat)
    pg, dg = og.get_dofmap()
    blobs = [d.halo_export()[0] for d in devs]
    for r in range(R):
        devs[r].halo_import(r - 1 if r > 0 else -1, blobs[r - 1] if r > 0 else None, r + 1 if r < R - 1 else -1, blobs[r + 1] if r < R - 1 else None)
    gfull, owned = [], []
    seen = np.zeros(Pg.ndof, dtype=int)
    for r in range(R):
        pl, dl = devs[r].get_dofmap()
        g = parallel.slab_global_dofs(pl, dl, Ps[r].ndof, slabs[r], pg, dg)
        assert (g >= 0).all()
        f, _ = devs[r].get_ownership(None)
        assert int(f.sum()) == devs[r].owned_rows()[0] == sum(int(v) for v in devs[r].owned_counts())
        gfull.append(g); owned.append(f.astype(bool))
        seen[g[owned[r]]] += 1
    assert (seen == 1).all()                   # every global DOF is owned by exactly one rank
    for r in range(R):
        d, P = devs[r], Ps[r]
        u = np.zeros(P.ndof)
        d.interpolate(P.sps, P.gfn, u)
        assert np.allclose(u, ug[gfull[r]], rtol=1e-14, atol=0)
        assert np.array_equal(d.get_constraints()[owned[r]], og.get_constraints()[gfull[r][owned[r]]])
        u = np.ascontiguousarray(ug[gfull[r]])
        d.jacobian(P.go, P.mat, u, overwrite=True)
        rp, ci = d.matrix_get_pattern(P.mat)
        va = d.matrix_get_values(P.mat)
        scale = np.abs(Ag.data).max()
        for row in np.where(owned[r])[0]:
            a, b = rp[row], rp[row + 1]
            ga, gb = Ag.indptr[gfull[r][row]], Ag.indptr[gfull[r][row] + 1]
            cols = gfull[r][ci[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb])
            assert np.abs(va[a:b][order] - Ag.data[ga:gb]).max() <= 1e-12 * scale
        res = np.zeros(P.ndof)
        d.residual(P.go, u, res)
        assert np.abs(res[owned[r]] - rg[gfull[r][owned[r]]]).max() <= 1e-12 * np.abs(rg).max()
    rng = np.random.default_rng(0)
    xg = rng.uniform(-1, 1, Pg.ndof)
    xd = [torch.from_numpy(np.where(owned[r], xg[gfull[r]], 1e300)).cuda() for r in range(R)]
    torch.cuda.synchronize()
    for r in range(R):
        devs[r].halo_push(xd[r])
        devs[r].synchronize()
    for r in range(R):
        devs[r].halo_wait(xd[r])
        devs[r].synchronize()
        assert np.array_equal(xd[r].cpu().numpy(), xg[gfull[r]])
    for d in devs:
        d.close()


MORTAR_CASES = [((4, 5, 6), 1, (capi.FE_Q1, capi.FE_Q1)), ((6, 8), 1, (capi.FE_Q1, capi.FE_Q1)), ((3, 4, 6), 2, (capi.FE_Q1, capi.FE_Q1)),
                ((3, 3, 6), 1, (capi.FE_Q2, capi.FE_Q1))]


def _global_oracle_mortar(n, ncomp, fe):
    import orc
    import problems
    og = orc.Oracle()
    Pg = problems.testmortarpoisson(og, n, ncomp=ncomp, fe=fe)
    ug = np.random.default_rng(5).uniform(-1, 1, Pg.ndof)
    og.jacobian(Pg.go, Pg.mat, ug, overwrite=True)
    rg = np.zeros(Pg.ndof)
    og.residual(Pg.go, ug, rg)
    return og, Pg, ug, rg


@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("n,ncomp,fe", MORTAR_CASES)
def test_two_contexts_on_one_gpu_mortar(n, ncomp, fe, R):
    """Enriched (mortar) couplings on a slab partition (BASELINE configs[3]): the coupling space is vertex-attached, so its block is
    owned by lattice planes like a Q1 block.  Two contexts of one process act as two ranks; owned rows (element rows with their
    mortar links, mortar rows) against the oracle's global matrix, residual, halo exchange of all blocks."""
    import torch
    import problems
    devs = [pkg.Mdb(0) for _ in range(R)]
    slabs = [parallel.slab(n, r, R) for r in range(R)]
    Ps = [problems.testmortarpoisson(devs[r], n, ncomp=ncomp, fe=fe, slab=slabs[r]) for r in range(R)]
    og, Pg, ug, rg = _global_oracle_mortar(n, ncomp, fe)
    Ag = og.csr(Pg.mat)
    blobs = [d.halo_export()[0] for d in devs]
    for r in range(R):
        devs[r].halo_import(r - 1 if r > 0 else -1, blobs[r - 1] if r > 0 else None, r + 1 if r < R - 1 else -1, blobs[r + 1] if r < R - 1 else None)
    # no cell carries the mortar DOFs, so the cell maps cannot give the local -> global map; every block of this system that is cut
    # is a Q1 block (lexicographic, cut axis slowest), except a Q2 element block, which goes through the cell maps
    counts = [[int(v) for v in d.owned_counts()] for d in devs]
    assert sum(sum(c) for c in counts) == Pg.ndof
    pg, dg = og.get_dofmap()
    gfull, owned = [], []
    for r in range(R):
        f, _ = devs[r].get_ownership(None)
        ow = f.astype(bool)
        off = devs[r].get_child_offsets()
        goff = og.get_child_offsets()
        g = -np.ones(Ps[r].ndof, dtype=np.int64)
        pl, dl = devs[r].get_dofmap()
        gc = parallel.slab_global_dofs(pl, dl, Ps[r].ndof, slabs[r], pg, dg)        # element blocks (-1 on the mortar blocks)
        for k in range(len(off) - 1):
            a, b = off[k], off[k + 1]
            if (gc[a:b] >= 0).all():
                g[a:b] = gc[a:b]
            else:                                                                   # mortar block: rank offsets
                below = sum(counts[q][k] for q in range(r))
                idx = np.nonzero(ow[a:b])[0]
                g[a + idx] = goff[k] + below + np.arange(idx.size)
        gfull.append(g); owned.append(ow)
    xs = [torch.from_numpy(np.where(owned[r], gfull[r], -1).astype(np.float64)).cuda() for r in range(R)]
    torch.cuda.synchronize()
    for r in range(R):
        devs[r].halo_push(xs[r])
    for r in range(R):
        devs[r].synchronize()
    for r in range(R):
        devs[r].halo_wait(xs[r])
        devs[r].synchronize()
        g = xs[r].cpu().numpy().astype(np.int64)
        assert (g >= 0).all(), "a ghost DOF was not covered by the halo exchange"
        assert np.array_equal(g[gfull[r] >= 0], gfull[r][gfull[r] >= 0])           # the two maps agree where both exist
        gfull[r] = g
    seen = np.zeros(Pg.ndof, dtype=int)
    for r in range(R):
        seen[gfull[r][owned[r]]] += 1
    assert (seen == 1).all()
    for r in range(R):
        d, P = devs[r], Ps[r]
        assert np.array_equal(d.get_constraints()[owned[r]], og.get_constraints()[gfull[r][owned[r]]])
        u = np.ascontiguousarray(ug[gfull[r]])
        d.jacobian(P.go, P.mat, u, overwrite=True)
        rp, ci = d.matrix_get_pattern(P.mat)
        va = d.matrix_get_values(P.mat)
        scale = np.abs(Ag.data).max()
        for row in np.where(owned[r])[0]:
            a, b = rp[row], rp[row + 1]
            ga, gb = Ag.indptr[gfull[r][row]], Ag.indptr[gfull[r][row] + 1]
            cols = gfull[r][ci[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb])
            assert np.abs(va[a:b][order] - Ag.data[ga:gb]).max() <= 1e-12 * scale
        res = np.zeros(P.ndof)
        d.residual(P.go, u, res)
        assert np.abs(res[owned[r]] - rg[gfull[r][owned[r]]]).max() <= 1e-12 * np.abs(rg).max()
    for d in devs:
        d.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    import problems
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        dev = pkg.Mdb(rank)
        s = parallel.slab(n, rank, world)
        P = problems.testpoisson(dev, n, slab=s)
        link = parallel.connect(dev, dist, rank, world)
        flags, gidx = dev.get_ownership(link.child_global_base)
        owned = flags.astype(bool)
        u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda")
        r = torch.zeros_like(u)
        z = torch.zeros_like(u)
        torch.cuda.synchronize()
        dev.interpolate(P.sps, P.gfn, u)
        # linearise at the ORACLE's interpolated u (device exp() differs from libm in the last bit, and the FD coupling
        # Jacobian amplifies one ulp of x to 1e-8): owned entries from the global vector, ghosts by a halo exchange
        og, Pg, ug, rg = _global_oracle(n)
        uh = u.cpu().numpy()
        uh[owned] = ug[gidx[owned]]
        u.copy_(torch.from_numpy(uh))
        torch.cuda.synchronize()
        dev.halo_push(u)
        dev.halo_wait(u)
        dev.jacobian(P.go, P.mat, u, overwrite=True)
        dev.residual(P.go, u, r)
        out = {}
        for name, method in (("cg", capi.SOLVER_CG), ("bicgstab", capi.SOLVER_BICGSTAB)):
            st = dev.solve(P.mat, r, z, method=method, precond=capi.PRECOND_JACOBI, reduction=1e-12, maxit=4000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        # block-Jacobi SeqSSOR / SeqILU0 over the ranks: each rank applies the sequential preconditioner of its
        # owned x owned block.  One CG step from x = 0 gives x_1 = lambda * M^-1 r (lambda a global scalar): its direction is
        # checked in the parent against a NumPy application on the extracted blocks; the full solves must reach the global solution.
        rp, ci = dev.matrix_get_pattern(P.mat)
        pre = {"rowptr": rp, "colidx": ci, "values": dev.matrix_get_values(P.mat), "owned": owned, "r": r.cpu().numpy()}
        for name, precond, sweeps in (("ssor", capi.PRECOND_SSOR, 2), ("ilu0", capi.PRECOND_ILU0, 0)):
            if sweeps:
                dev.set_ssor_sweeps(sweeps)
            dev.solve(P.mat, r, z, method=capi.SOLVER_CG, precond=precond, fixed_iters=1)
            pre[name] = z.cpu().numpy().copy()
        dev.set_ssor_sweeps(3)
        for name, method, precond in (("bcgs_ssor", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR), ("cg_ilu0", capi.SOLVER_CG, capi.PRECOND_ILU0)):
            st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-12, maxit=2000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        # distributed SpMV through the public entry point
        xg = np.random.default_rng(0).uniform(-1, 1, link.global_ndof)
        x = torch.from_numpy(np.where(owned, xg[np.maximum(gidx, 0)], 1e300)).cuda()
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        dev.spmv(P.mat, x, y)
        dev.synchronize()
        import os
        assert (dev.scalars_exchanges() > 0) == (not os.environ.get("MDB_SCALARS_NCCL")), "the Krylov scalars did not go through the peer windows"
        ret[rank] = {"gidx": gidx[owned], "sol": out, "y": y.cpu().numpy()[owned], "ndof": link.global_ndof, "pre": pre}
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [(6, 6, 12)])
def test_multi_gpu_solve_matches_global_oracle(n):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    og, Pg, ug, rg = _global_oracle(n)
    Ag = og.csr(Pg.mat)
    zg = spla.spsolve(sp.csc_matrix(Ag), rg)
    yg = Ag @ np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
    seen = np.zeros(Pg.ndof, dtype=int)
    iters = set()
    for rank in range(world):
        o = ret[rank]
        assert o["ndof"] == Pg.ndof
        seen[o["gidx"]] += 1
        for name in ("cg", "bicgstab", "bcgs_ssor", "cg_ilu0"):
            it, zz = o["sol"][name]
            assert np.abs(zz - zg[o["gidx"]]).max() <= 1e-10 * np.abs(zg).max(), name
        iters.add(o["sol"]["cg"][0])
        assert np.abs(o["y"] - yg[o["gidx"]]).max() <= 1e-12 * np.abs(yg).max()
    assert (seen == 1).all()          # every global DOF is owned by exactly one rank
    assert len(iters) == 1
    for name in ("bcgs_ssor", "cg_ilu0"):
        assert len({ret[rank]["sol"][name][0] for rank in range(world)}) == 1     # same iteration count on every rank
    # block-Jacobi preconditioners: reference application on the owned x owned blocks
    for name, sweeps in (("ssor", 2), ("ilu0", 0)):
        dev_dir, ref_dir = [], []
        for rank in range(world):
            pre = ret[rank]["pre"]
            ow = pre["owned"]
            A = sp.csr_matrix((pre["values"], pre["colidx"], pre["rowptr"]), shape=(len(ow), len(ow)))[ow][:, ow].tocsr()
            A.sort_indices()
            d = pre["r"][ow]
            v = np.zeros(len(d))
            if name == "ssor":                              # SeqSSOR(A, sweeps, 1.0): forward then backward Gauss-Seidel sweeps
                for _ in range(sweeps):
                    for order in (range(len(d)), range(len(d) - 1, -1, -1)):
                        for i in order:
                            a, b = A.indptr[i], A.indptr[i + 1]
                            v[i] += (d[i] - A.data[a:b] @ v[A.indices[a:b]]) / A[i, i]
            else:                                           # SeqILU0: ILU(0) on the block's pattern, then the two triangular solves
                Ad = A.toarray()
                n_ = len(d)
                pat = np.zeros((n_, n_), dtype=bool)        # the STRUCTURAL pattern (stored zeros count: Dirichlet rows, the
                for i in range(n_):                         # exact-zero edge entries of the trilinear stiffness matrix)
                    pat[i, A.indices[A.indptr[i]:A.indptr[i + 1]]] = True
                for i in range(n_):
                    for k in np.nonzero(pat[i, :i])[0]:
                        Ad[i, k] = Ad[i, k] / Ad[k, k]
                        js = np.nonzero(pat[i, k + 1:] & pat[k, k + 1:])[0] + k + 1
                        Ad[i, js] -= Ad[i, k] * Ad[k, js]
                y_ = np.zeros(n_)
                for i in range(n_):
                    y_[i] = d[i] - Ad[i, :i] @ y_[:i]
                for i in range(n_ - 1, -1, -1):
                    v[i] = (y_[i] - Ad[i, i + 1:] @ v[i + 1:]) / Ad[i, i]
            ref_dir.append(v)
            dev_dir.append(pre[name][ow])
        dev_all, ref_all = np.concatenate(dev_dir), np.concatenate(ref_dir)
        err = float(np.abs(dev_all / np.abs(dev_all).max() - ref_all / np.abs(ref_all).max()).max())
        ratios = [float(np.abs(a).max() / np.abs(b).max()) for a, b in zip(dev_dir, ref_dir)]
        assert err <= 1e-13, "%s: err %.3e, per-rank max ratios %s" % (name, err, ratios)


def _rank_main_qk(rank, world, port, n, fe, ret):
    import torch
    import torch.distributed as dist
    import problems
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        dev = pkg.Mdb(rank)
        s = parallel.slab(n, rank, world)
        P = problems.testpoisson(dev, n, fe=fe, slab=s)
        link = parallel.connect(dev, dist, rank, world)
        og, Pg, ug, rg = _global_oracle_fe(n, fe)
        pg, dg = og.get_dofmap()
        pl, dl = dev.get_dofmap()
        gidx = parallel.slab_global_dofs(pl, dl, P.ndof, s, pg, dg)
        owned = dev.get_ownership(None)[0].astype(bool)
        assert link.global_ndof == Pg.ndof
        u = torch.from_numpy(np.where(owned, ug[gidx], 1e300)).cuda()
        r = torch.zeros_like(u)
        z = torch.zeros_like(u)
        torch.cuda.synchronize()
        dev.halo_push(u)
        dev.halo_wait(u)
        dev.synchronize()
        halo_ok = bool(np.array_equal(u.cpu().numpy(), ug[gidx]))
        dev.jacobian(P.go, P.mat, u, overwrite=True)
        dev.residual(P.go, u, r)
        out = {}
        for name, method, precond in (("cg", capi.SOLVER_CG, capi.PRECOND_JACOBI), ("bicgstab", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI),
                                      ("bcgs_ssor", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR)):
            st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-12, maxit=4000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        xg = np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
        x = torch.from_numpy(np.where(owned, xg[gidx], 1e300)).cuda()
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        dev.spmv(P.mat, x, y)
        dev.synchronize()
        ret[rank] = {"gidx": gidx[owned], "sol": out, "y": y.cpu().numpy()[owned], "res": r.cpu().numpy()[owned], "halo_ok": halo_ok}
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fe,n", [((capi.FE_Q1, capi.FE_Q2), (4, 4, 8)), ((capi.FE_Q2, capi.FE_Q2), (3, 4, 8))])
def test_multi_gpu_qk_solve_matches_global_oracle(fe, n):
    """Qk (k = 2) child spaces on a slab partition, one process per GPU: residual, distributed SpMV and the Krylov solves (Jacobi,
    block-Jacobi SSOR) against the oracle's global system."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_rank_main_qk, args=(world, _free_port(), n, fe, ret), nprocs=world, join=True)
    og, Pg, ug, rg = _global_oracle_fe(n, fe)
    Ag = og.csr(Pg.mat)
    zg = spla.spsolve(sp.csc_matrix(Ag), rg)
    yg = Ag @ np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
    seen = np.zeros(Pg.ndof, dtype=int)
    for rank in range(world):
        o = ret[rank]
        assert o["halo_ok"]
        seen[o["gidx"]] += 1
        assert np.abs(o["res"] - rg[o["gidx"]]).max() <= 1e-12 * np.abs(rg).max()
        for name in ("cg", "bicgstab", "bcgs_ssor"):
            it, zz = o["sol"][name]
            assert np.abs(zz - zg[o["gidx"]]).max() <= 1e-10 * np.abs(zg).max(), name
        assert np.abs(o["y"] - yg[o["gidx"]]).max() <= 1e-12 * np.abs(yg).max()
    assert (seen == 1).all()
    for name in ("cg", "bicgstab", "bcgs_ssor"):
        assert len({ret[rank]["sol"][name][0] for rank in range(world)}) == 1


def _dg_global(n, k):
    import orc
    import problems
    og = orc.Oracle()
    Pg = problems.dirichlet_neumann_dg(og, n, k=k)
    xg = np.random.default_rng(3).uniform(-1, 1, Pg.ndof)
    og.jacobian(Pg.go, Pg.mat, xg, overwrite=True)
    rg = np.zeros(Pg.ndof)
    og.residual(Pg.go, xg, rg)
    return og, Pg, xg, rg


def _rank_main_dg(rank, world, port, n, k, ret):
    import torch
    import torch.distributed as dist
    import problems
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        dev = pkg.Mdb(rank)
        s = parallel.slab2(n, rank, world)
        P = problems.dirichlet_neumann_dg(dev, n, k=k, box=s)
        pl, dl = dev.get_dofmap()
        ids, owner = parallel.cell_blocked_ids(pl, dl, P.ndof, s)
        link = parallel.connect_lists(dev, dist, rank, world, ids, owner)
        owned = owner == rank
        og, Pg, xg, rg = _dg_global(n, k)
        pg, dg = og.get_dofmap()
        gidx = parallel.slab_global_dofs(pl, dl, P.ndof, s, pg, dg)
        u = torch.from_numpy(np.where(owned, xg[gidx], 1e300)).cuda()
        r = torch.zeros_like(u)
        z = torch.zeros_like(u)
        torch.cuda.synchronize()
        dev.halo_push(u)
        dev.halo_wait(u)
        dev.synchronize()
        halo_ok = bool(np.array_equal(u.cpu().numpy(), xg[gidx]))
        dev.jacobian(P.go, P.mat, u, overwrite=True)
        dev.residual(P.go, u, r)
        rp, ci = dev.matrix_get_pattern(P.mat)
        va = dev.matrix_get_values(P.mat)
        out = {}
        for name, method, precond in (("bicgstab", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI), ("bcgs_ssor", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR)):
            st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-12, maxit=8000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        xr = np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
        x = torch.from_numpy(np.where(owned, xr[gidx], 1e300)).cuda()
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        dev.spmv(P.mat, x, y)
        dev.synchronize()
        ret[rank] = dict(gid=gidx, owned=owned, rowptr=rp, colidx=ci, values=va, res=r.cpu().numpy(), sol=out, y=y.cpu().numpy()[owned],
                         halo_ok=halo_ok, neighbours=link.neighbours)
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [((8, 12), 2), ((4, 3, 8), 1)])
def test_multi_gpu_dg_partition_matches_global_oracle(n, k):
    """QkDG spaces + ConvectionDiffusionDG + the CVCF coupling (BASELINE configs[0] as shipped) on a slab partition with ghost cell
    layers on both sides and index-list halos: the owned rows (pattern, values, residual) equal the global oracle's without matrix
    communication; halo exchange, SpMV and BiCGSTAB (Jacobi, block-Jacobi SSOR) reproduce the global results."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_rank_main_dg, args=(world, _free_port(), n, k, ret), nprocs=world, join=True)
    og, Pg, xg, rg = _dg_global(n, k)
    Ag = og.csr(Pg.mat)
    zg = spla.spsolve(sp.csc_matrix(Ag), rg)
    yg = Ag @ np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
    seen = np.zeros(Pg.ndof, dtype=int)
    scale = np.abs(Ag.data).max()
    rscale = abs(Ag) @ np.abs(xg) + np.abs(rg)
    for rank in range(world):
        o = ret[rank]
        assert o["halo_ok"]
        gid, owned = o["gid"], o["owned"]
        seen[gid[owned]] += 1
        rp, ci, va = o["rowptr"], o["colidx"], o["values"]
        for row in np.where(owned)[0]:
            a, b = rp[row], rp[row + 1]
            ga, gb = Ag.indptr[gid[row]], Ag.indptr[gid[row] + 1]
            cols = gid[ci[a:b]]
            order = np.argsort(cols)
            assert np.array_equal(cols[order], Ag.indices[ga:gb])
            assert np.abs(va[a:b][order] - Ag.data[ga:gb]).max() <= 1e-12 * scale
        assert np.all(np.abs(o["res"][owned] - rg[gid[owned]]) <= 1e-12 * np.maximum(rscale[gid[owned]], 1e-3 * np.abs(rg).max()))
        for name in ("bicgstab", "bcgs_ssor"):
            it, zz = o["sol"][name]
            assert np.abs(zz - zg[gid[owned]]).max() <= 1e-10 * np.abs(zg).max(), name
        assert np.abs(o["y"] - yg[gid[owned]]).max() <= 1e-12 * np.abs(yg).max()
    assert (seen == 1).all()
    for name in ("bicgstab", "bcgs_ssor"):
        assert len({ret[rank]["sol"][name][0] for rank in range(world)}) == 1


def _rank_main_mortar(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    import problems
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        dev = pkg.Mdb(rank)
        s = parallel.slab(n, rank, world)
        P = problems.testmortarpoisson(dev, n, slab=s)
        link = parallel.connect(dev, dist, rank, world)
        flags, gidx = dev.get_ownership(link.child_global_base)          # all three blocks are Q1-like (vertex-attached, lexicographic)
        owned = flags.astype(bool)
        og, Pg, ug, rg = _global_oracle_mortar(n, 1, (capi.FE_Q1, capi.FE_Q1))
        assert link.global_ndof == Pg.ndof
        u = torch.from_numpy(np.where(owned, ug[np.maximum(gidx, 0)], 1e300)).cuda()
        r = torch.zeros_like(u)
        z = torch.zeros_like(u)
        torch.cuda.synchronize()
        dev.halo_push(u)
        dev.halo_wait(u)
        dev.jacobian(P.go, P.mat, u, overwrite=True)
        dev.residual(P.go, u, r)
        dev.set_ssor_sweeps(3)
        st = dev.solve(P.mat, r, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-12, maxit=3000)
        sol = (bool(st.converged), st.iterations, z.cpu().numpy()[owned].copy())
        xg = np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
        x = torch.from_numpy(np.where(owned, xg[np.maximum(gidx, 0)], 1e300)).cuda()
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        dev.spmv(P.mat, x, y)
        dev.synchronize()
        ret[rank] = {"gidx": gidx[owned], "sol": sol, "y": y.cpu().numpy()[owned], "res": r.cpu().numpy()[owned]}
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [(4, 4, 8)])
def test_multi_gpu_mortar_matches_global_oracle(n):
    """testmortarpoisson (enriched coupling, Q1 mortar space) on a slab partition, one process per GPU: residual, distributed SpMV and
    the reference's solver configuration (BiCGSTAB + SeqSSOR(3), block-Jacobi over the ranks) against the oracle's global system."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_rank_main_mortar, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    og, Pg, ug, rg = _global_oracle_mortar(n, 1, (capi.FE_Q1, capi.FE_Q1))
    Ag = og.csr(Pg.mat)
    zg = spla.spsolve(sp.csc_matrix(Ag), rg)
    yg = Ag @ np.random.default_rng(0).uniform(-1, 1, Pg.ndof)
    seen = np.zeros(Pg.ndof, dtype=int)
    for rank in range(world):
        o = ret[rank]
        seen[o["gidx"]] += 1
        assert np.abs(o["res"] - rg[o["gidx"]]).max() <= 1e-12 * np.abs(rg).max()
        assert np.abs(o["y"] - yg[o["gidx"]]).max() <= 1e-12 * np.abs(yg).max()
        conv, it, zz = o["sol"]
        assert conv, "BiCGSTAB + block-Jacobi SSOR(3) did not converge on the mortar system"
        assert np.abs(zz - zg[o["gidx"]]).max() <= 1e-9 * np.abs(zg).max()
    assert (seen == 1).all()
    assert len({ret[rank]["sol"][1] for rank in range(world)}) == 1


# ---- unstructured partition (recursive coordinate bisection, index-list halos: mdb_partition_lists) ------------------------------
def _um_global(nx, ny, seed, cvcf):
    import um_problems as ump
    xy, conn = ump.square_mesh(nx, ny, seed=seed)
    sd = ump.halves(xy, conn)
    og = ump.oracle()
    Hg = ump.setup(og, xy, conn, sd, cvcf=cvcf, jac_mode=capi.JAC_ANALYTIC)
    og.jacobian(Hg["go"], Hg["mat"], Hg["x0"], overwrite=True)
    rg = og.residual(Hg["go"], Hg["x0"], np.zeros(Hg["ndof"]))
    return xy, conn, sd, og, Hg, rg


def _um_rank_main(rank, world, port, mesh, ret):
    import torch
    import torch.distributed as dist
    import um_problems as ump
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny, seed, cvcf = mesh
        xy, conn, sd, og, Hg, rg = _um_global(nx, ny, seed, cvcf)
        R = parallel.um_rank_problem(xy, conn, sd, world, rank, [0, 1])
        dev = pkg.Mdb(rank)
        H = ump.setup(dev, R.xy, R.conn, R.sd, cvcf=cvcf, jac_mode=capi.JAC_ANALYTIC)
        assert H["ndof"] == R.ndof
        link = parallel.um_connect(dev, dist, rank, world, R.global_dof, R.owner)
        owned = R.owned
        # the same linearisation point as the global problem: owned entries from the global vector, ghosts by a halo exchange
        u = torch.from_numpy(np.where(owned, Hg["x0"][R.global_dof], 1e300)).cuda()
        torch.cuda.synchronize()
        dev.halo_push(u); dev.halo_wait(u); dev.synchronize()
        halo_ok = bool(np.array_equal(u.cpu().numpy(), Hg["x0"][R.global_dof]))
        r = torch.zeros_like(u); z = torch.zeros_like(u)
        dev.jacobian(H["go"], H["mat"], u, overwrite=True)
        dev.residual(H["go"], u, r)
        rp, ci = dev.matrix_get_pattern(H["mat"])
        va = dev.matrix_get_values(H["mat"])
        out = {}
        for name, method, precond in (("cg", capi.SOLVER_CG, capi.PRECOND_JACOBI), ("bicgstab", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI),
                                      ("bcgs_ssor", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR)):
            st = dev.solve(H["mat"], r, z, method=method, precond=precond, reduction=1e-12, maxit=4000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        xg = np.random.default_rng(0).uniform(-1, 1, R.global_ndof)
        x = torch.from_numpy(np.where(owned, xg[R.global_dof], 1e300)).cuda()
        y = torch.zeros_like(x)
        torch.cuda.synchronize()
        dev.spmv(H["mat"], x, y)
        dev.synchronize()
        ret[rank] = dict(gid=R.global_dof, owned=owned, rowptr=rp, colidx=ci, values=va, res=r.cpu().numpy(), sol=out, y=y.cpu().numpy()[owned], halo_ok=halo_ok,
                         neighbours=link.neighbours, owned_rows=dev.owned_rows()[0])
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mesh", [(14, 11, 5, False), (12, 12, 8, True)])
def test_multi_gpu_unstructured_partition_matches_global_oracle(mesh):
    """Recursive coordinate bisection over the GPUs of the box: the owned rows of every rank (pattern, values, residual) equal the
    global oracle's without matrix communication; halo exchange, SpMV and the Krylov solves (Jacobi, block-Jacobi SSOR) reproduce
    the global results."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_um_rank_main, args=(world, _free_port(), mesh, ret), nprocs=world, join=True)
    xy, conn, sd, og, Hg, rg = _um_global(*mesh)
    rpg, cig = og.matrix_get_pattern(Hg["mat"])
    Ag = sp.csr_matrix((og.matrix_get_values(Hg["mat"]), cig, rpg), shape=(Hg["ndof"],) * 2)
    zg = spla.spsolve(Ag.tocsc(), rg)
    yg = Ag @ np.random.default_rng(0).uniform(-1, 1, Hg["ndof"])
    seen = np.zeros(Hg["ndof"], dtype=int)
    scale = np.abs(Ag.data).max()
    for rank in range(world):
        o = ret[rank]
        assert o["halo_ok"] and o["owned_rows"] == int(o["owned"].sum())
        gid, owned = o["gid"], o["owned"]
        seen[gid[owned]] += 1
        for row in np.nonzero(owned)[0]:
            a, b = o["rowptr"][row], o["rowptr"][row + 1]
            ga, gb = Ag.indptr[gid[row]], Ag.indptr[gid[row] + 1]
            cols = gid[o["colidx"][a:b]]
            assert np.array_equal(cols, Ag.indices[ga:gb])                           # monotone local numbering: same order, bit-exact pattern
            assert np.abs(o["values"][a:b] - Ag.data[ga:gb]).max() <= 1e-12 * scale
        assert np.abs(o["res"][owned] - rg[gid[owned]]).max() <= 1e-12 * np.abs(rg).max()
        for name in ("cg", "bicgstab", "bcgs_ssor"):
            assert np.abs(o["sol"][name][1] - zg[gid[owned]]).max() <= 1e-10 * np.abs(zg).max(), name
        assert np.abs(o["y"] - yg[gid[owned]]).max() <= 1e-12 * np.abs(yg).max()
    assert (seen == 1).all()
    assert len({ret[r]["sol"]["cg"][0] for r in range(world)}) == 1


# ---- multigrid on a slab partition (MDB_PRECOND_MG with mdb_mg_setup / mdb_mg_level) ---------------------------------------------------
def _mg_rank_main(rank, world, port, n, ret):
    import torch
    import torch.distributed as dist
    import problems
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        dev = pkg.Mdb(rank)
        P = problems.testpoisson(dev, n, slab=parallel.slab(n, rank, world))
        link = parallel.connect(dev, dist, rank, world)
        nlev = parallel.connect_mg(dev, P.mat, dist, rank, world)
        flags, gidx = dev.get_ownership(link.child_global_base)
        owned = flags.astype(bool)
        og, Pg, ug, rg = _global_oracle(n)
        u = torch.from_numpy(np.where(owned, ug[np.maximum(gidx, 0)], 0.0)).cuda()
        r = torch.zeros_like(u); z = torch.zeros_like(u)
        torch.cuda.synchronize()
        dev.halo_push(u); dev.halo_wait(u)
        dev.jacobian(P.go, P.mat, u, overwrite=True)
        dev.residual(P.go, u, r)
        out = {}
        for name, method, precond in (("cg_jacobi", capi.SOLVER_CG, capi.PRECOND_JACOBI), ("cg_mg", capi.SOLVER_CG, capi.PRECOND_MG), ("bcgs_mg", capi.SOLVER_BICGSTAB, capi.PRECOND_MG)):
            st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-13, maxit=2000)
            assert st.converged, name
            out[name] = (st.iterations, z.cpu().numpy()[owned].copy())
        ret[rank] = {"gidx": gidx[owned], "sol": out, "levels": nlev}
        dist.barrier()
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [(8, 8, 16), (16, 16, 32)])
def test_multi_gpu_multigrid_matches_global_oracle(n):
    """The V-cycle on a slab partition: every rank coarsens its slab, the coarse levels exchange halos through their own windows
    and reduce over the shared communicator; the preconditioned solves reach the global oracle's solution in a few iterations."""
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_mg_rank_main, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    og, Pg, ug, rg = _global_oracle(n)
    zg = spla.spsolve(sp.csc_matrix(og.csr(Pg.mat)), rg)
    for rank in range(world):
        o = ret[rank]
        # coarsening needs four owned layers per rank (a coarse slab keeps two): 16 / 32 layers give 3 / 4 levels on 2 ranks, 2 / 3 on 4
        assert o["levels"] >= (3 if n[2] // world >= 8 else 2)
        for name in ("cg_jacobi", "cg_mg", "bcgs_mg"):
            assert np.abs(o["sol"][name][1] - zg[o["gidx"]]).max() <= 1e-10 * np.abs(zg).max(), name
        # a two-level hierarchy (4 layers per rank) ends on a Chebyshev-smoothed coarse level, not on a near-exact one: fewer iterations, no factor
        gain = 2 if o["levels"] >= 3 else 1
        assert o["sol"]["cg_mg"][0] * gain < o["sol"]["cg_jacobi"][0], (o["sol"]["cg_mg"][0], o["sol"]["cg_jacobi"][0])
    assert len({ret[r]["sol"]["cg_mg"][0] for r in range(world)}) == 1
