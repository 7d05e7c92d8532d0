"""MDB_PRECOND_MG (csrc/mg.cu): the geometric multigrid V-cycle as a preconditioner of CG / BiCGSTAB.  The preconditioner is not in the
reference (its backends use SeqSSOR / SeqILU0); what must hold is that it changes only the iteration count: the solution of the same
assembled system agrees with the oracle's to 1e-10, with far fewer iterations than Jacobi and an iteration count that does not grow
with the mesh."""
import importlib

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import orc
import problems

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi


def _assemble(api, P, u=None):
    if u is None:
        u = np.zeros(P.ndof)
        api.interpolate(P.sps, P.gfn, u)
    api.jacobian(P.go, P.mat, u, overwrite=True)
    r = np.zeros(P.ndof)
    api.residual(P.go, u, r)
    return u, r


@pytest.mark.parametrize("n,kw", [((32, 32), {}), ((16, 16, 16), {}), ((24, 16, 8), dict(split_axis=0, intensity=10.0)),
                                  ((32, 48), dict(coupling=capi.OP_PROPFLOW_COUPLING, intensity=3.0))])
def test_cg_with_multigrid_matches_oracle_solution(n, kw):
    dev, ora = pkg.Mdb(0), orc.Oracle()
    Pd, Po = problems.testpoisson(dev, n, **kw), problems.testpoisson(ora, n, **kw)
    uo, ro = _assemble(ora, Po)
    ud, rd = _assemble(dev, Pd, uo)          # the oracle's linearisation point: device exp() differs in the last bit and the FD coupling Jacobian amplifies it
    zo = spla.spsolve(sp.csc_matrix(ora.csr(Po.mat)), ro)              # the ORACLE's system, solved directly
    z = np.zeros(Pd.ndof)
    st = dev.solve(Pd.mat, rd, z, method=capi.SOLVER_CG, precond=capi.PRECOND_MG, reduction=1e-14, maxit=200)
    assert st.converged and dev.matrix_mg_levels(Pd.mat) >= 3
    assert np.abs(z - zo).max() <= 1e-10 * np.abs(zo).max()
    zj = np.zeros(Pd.ndof)
    stj = dev.solve(Pd.mat, rd, zj, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-14, maxit=5000)
    assert stj.converged and st.iterations * 3 < stj.iterations, (st.iterations, stj.iterations)
    # BiCGSTAB takes the V-cycle as a right preconditioner
    zb = np.zeros(Pd.ndof)
    stb = dev.solve(Pd.mat, rd, zb, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_MG, reduction=1e-14, maxit=200)
    assert stb.converged and np.abs(zb - zo).max() <= 1e-10 * np.abs(zo).max()
    dev.close(); ora.close()


def test_iteration_count_does_not_grow_with_the_mesh():
    its = []
    for n in ((32, 32, 32), (64, 64, 64)):
        dev = pkg.Mdb(0)
        P = problems.testpoisson(dev, n)
        u, r = _assemble(dev, P)
        z = np.zeros(P.ndof)
        st = dev.solve(P.mat, r, z, method=capi.SOLVER_CG, precond=capi.PRECOND_MG, reduction=1e-10, maxit=200)
        assert st.converged
        its.append(st.iterations)
        dev.close()
    assert its[1] <= its[0] + 3 and its[1] < 40, its


def test_one_step_matrix_coarsens_through_the_assembly_log():
    """J = J1 + dt J0 (two mdb_jacobian calls with weights): the coarse levels replay the same calls; implicit Euler steps with the
    multigrid-preconditioned solve agree with the Jacobi-preconditioned ones."""
    n = (32, 32)
    dev = pkg.Mdb(0)
    P = problems.instationary(dev, n)
    its = {}

    def solver(precond):
        def f(mat, R, z):
            st = dev.solve(mat, R, z, method=capi.SOLVER_BICGSTAB, precond=precond, reduction=1e-13, maxit=5000)
            assert st.converged
            its.setdefault(precond, []).append(st.iterations)
        return f
    u0 = np.zeros(P.ndof)
    dev.interpolate(P.sps, P.gfn, u0, time=0.0)
    ua, ub = u0.copy(), u0.copy()
    for step, dt in ((1, 0.04), (2, 0.04), (3, 0.01)):            # the third step changes dt: the log changes, the hierarchy is re-assembled
        t = 0.04 * min(step, 2) + (0.01 if step == 3 else 0.0)
        ua = problems.implicit_euler_step(dev, P, ua, t, dt, solver(capi.PRECOND_JACOBI))
        ub = problems.implicit_euler_step(dev, P, ub, t, dt, solver(capi.PRECOND_MG))
        assert np.abs(ua - ub).max() <= 1e-10 * max(np.abs(ua).max(), 1e-300), step
    assert max(its[capi.PRECOND_MG]) * 3 < max(its[capi.PRECOND_JACOBI])
    dev.close()


def test_refuses_what_it_cannot_coarsen():
    dev = pkg.Mdb(0)
    P = problems.testpoisson(dev, (8, 8), fe=(capi.FE_Q1, capi.FE_Q2))
    u, r = _assemble(dev, P)
    with pytest.raises(pkg.MdbError, match="Q1 child spaces only"):
        dev.solve(P.mat, r, np.zeros(P.ndof), method=capi.SOLVER_CG, precond=capi.PRECOND_MG)
    dev.close()


@pytest.mark.parametrize("n,k", [((32, 32), 2), ((32, 32), 1), ((8, 8, 8), 1)])
def test_multigrid_on_qkdg_spaces(n, k):
    """BASELINE configs[0] as shipped (DG-Qk, SIPG, CVCF coupling): nested spaces, the transfers evaluate the coarse cell's Lagrange
    basis at the fine nodes.  The system is that of the oracle; the V-cycle only changes the iteration count."""
    dev, ora = pkg.Mdb(0), orc.Oracle()
    Pd, Po = problems.dirichlet_neumann_dg(dev, n, k=k), problems.dirichlet_neumann_dg(ora, n, k=k)
    uo = np.zeros(Po.ndof)
    ora.jacobian(Po.go, Po.mat, uo, overwrite=True)
    dev.jacobian(Pd.go, Pd.mat, uo, overwrite=True)
    ro, rd = np.zeros(Po.ndof), np.zeros(Po.ndof)
    ora.residual(Po.go, uo, ro)
    dev.residual(Pd.go, uo, rd)
    zo = spla.spsolve(sp.csc_matrix(ora.csr(Po.mat)), ro)
    z = np.zeros(Pd.ndof)
    st = dev.solve(Pd.mat, rd, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_MG, reduction=1e-13, maxit=300)
    assert st.converged and dev.matrix_mg_levels(Pd.mat) >= 3
    assert np.abs(z - zo).max() <= 1e-9 * np.abs(zo).max()
    zj = np.zeros(Pd.ndof)
    stj = dev.solve(Pd.mat, rd, zj, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_JACOBI, reduction=1e-13, maxit=20000)
    assert stj.converged and st.iterations * 3 < stj.iterations, (st.iterations, stj.iterations)
    dev.close(); ora.close()
