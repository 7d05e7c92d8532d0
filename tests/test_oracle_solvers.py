"""Pins for the oracle's linear-solver stack (no GPU).  The device's SSOR / ILU0 / Krylov loops are checked against the
oracle (iteration counts, tests/test_gpu_parity.py); here the oracle itself is checked against implementations that share no code
with it: SciPy triangular solves for the Gauss-Seidel sweeps of SeqSSOR(A, n, 1.0), a dense textbook ILU(0) with its defining
property (L U = A on the pattern of A), and the textbook recurrences of preconditioned CG (Hestenes-Stiefel) and right-preconditioned
BiCGSTAB (van der Vorst) written with NumPy — compared iterate by iterate (maxit = 1, 2, 3, ...) and in the iteration counts.
[UPSTREAM] what ISTL computes is dune-istl's gsetc.hh (bsorf / bsorb), ilu.hh (bilu0_decomposition / bilu_backsolve) and
solvers.hh; this file pins the *mathematics* the oracle attributes to them."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import orc
import problems

capi = orc.capi


# ---- independent preconditioners ------------------------------------------------------------------------------------------------------
def jacobi(A):
    d = A.diagonal()
    return lambda r: r / d


def ssor(A, sweeps):
    """n symmetric Gauss-Seidel sweeps from v = 0 in matrix form: v += (D + L)^-1 (d - A v), then v += (D + U)^-1 (d - A v)."""
    DL = sp.tril(A, 0).tocsr()
    DU = sp.triu(A, 0).tocsr()

    def apply(d):
        v = np.zeros_like(d)
        for _ in range(sweeps):
            v = v + sla.spsolve_triangular(DL, d - A @ v, lower=True)
            v = v + sla.spsolve_triangular(DU, d - A @ v, lower=False)
        return v
    return apply


def ilu0_dense(A):
    """Textbook ILU(0) (Saad, Iterative Methods, Alg. 10.4, IKJ variant) on a dense copy; returns unit-lower L and upper U."""
    n = A.shape[0]
    A = A.tocsr()
    pat = np.zeros((n, n), dtype=bool)                   # the STORED positions (explicit zeros included: BCRS pattern, not the values)
    pat[np.repeat(np.arange(n), np.diff(A.indptr)), A.indices] = True
    a = A.toarray().astype(float)
    for i in range(1, n):
        for k in range(i):
            if not pat[i, k]:
                continue
            a[i, k] /= a[k, k]
            for j in range(k + 1, n):
                if pat[i, j]:
                    a[i, j] -= a[i, k] * a[k, j]
    L = np.tril(a, -1) + np.eye(n)
    U = np.triu(a)
    return L, U, pat


def ilu0(A):
    L, U, _ = ilu0_dense(A)
    Ls, Us = sp.csr_matrix(L), sp.csr_matrix(U)
    return lambda r: sla.spsolve_triangular(Us, sla.spsolve_triangular(Ls, r, lower=True, unit_diagonal=True), lower=False)


# ---- independent Krylov recurrences; every function returns the list of iterates x_1, x_2, ... and the defect norms -------------------------
def pcg(A, b, M, maxit, reduction):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z
    xs, defects = [], []
    d0 = np.linalg.norm(b)
    for _ in range(maxit):
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        xs.append(x.copy()); defects.append(np.linalg.norm(r))
        if defects[-1] <= reduction * d0:
            break
        z = M(r); rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return xs, defects


def bicgstab_right(A, b, M, maxit, reduction):
    """van der Vorst's BiCGSTAB with right preconditioning (y = M^-1 p, z = M^-1 s) and the convergence test after the half step
    as well as after the full step, which is where ISTL's BiCGSTABSolver counts (it reports the half step as iteration it - 0.5;
    the oracle and the device count it as `it`)."""
    x = np.zeros_like(b); r = b.copy(); rt = r.copy()
    rho = alpha = omega = 1.0
    v = np.zeros_like(b); p = np.zeros_like(b)
    xs, defects = [], []
    d0 = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        rho_new = rt @ r
        if it == 1:
            p = r.copy()
        else:
            p = r + (rho_new / rho) * (alpha / omega) * (p - omega * v)
        y = M(p); v = A @ y
        alpha = rho_new / (rt @ v)
        x = x + alpha * y
        s = r - alpha * v
        if np.linalg.norm(s) <= reduction * d0:
            xs.append(x.copy()); defects.append(np.linalg.norm(s))
            break
        z = M(s); t = A @ z
        omega = (t @ s) / (t @ t)
        x = x + omega * z
        r = s - omega * t
        rho = rho_new
        xs.append(x.copy()); defects.append(np.linalg.norm(r))
        if defects[-1] <= reduction * d0:
            break
    return xs, defects


# ---- systems -----------------------------------------------------------------------------------------------------------------------------
def _system(name):
    o = orc.Oracle()
    if name == "testpoisson_q1q2":                       # the shipped driver's spaces on a small mesh (unsymmetric at the FD-noise level)
        P = problems.testpoisson(o, (6, 6), fe=(capi.FE_Q1, capi.FE_Q2))
    elif name == "testpoisson_3d":
        P = problems.testpoisson(o, (3, 4, 3))
    else:
        P = problems.testpoisson(o, (8, 8), coupling=capi.OP_PROPFLOW_COUPLING)
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    r = np.zeros(P.ndof); o.residual(P.go, u, r)
    return o, P, o.csr(P.mat), r


SYSTEMS = ["testpoisson_q1q2", "testpoisson_3d", "propflow_2d"]
PRECONDS = [("jacobi", capi.PRECOND_JACOBI, 0), ("ssor1", capi.PRECOND_SSOR, 1), ("ssor3", capi.PRECOND_SSOR, 3), ("ilu0", capi.PRECOND_ILU0, 0)]


def _independent(A, name, sweeps):
    return {"jacobi": jacobi, "ilu0": ilu0}[name](A) if not name.startswith("ssor") else ssor(A, sweeps)


def test_textbook_ilu0_reproduces_the_matrix_on_its_pattern():
    """The defining property of ILU(0): (L U)_ij = a_ij wherever a_ij is stored.  A check of the independent implementation itself."""
    _, _, A, _ = _system("testpoisson_q1q2")
    L, U, pat = ilu0_dense(A)
    E = L @ U - A.toarray()
    assert np.abs(E[pat]).max() <= 1e-13 * np.abs(A.data).max()
    assert np.abs(E[~pat]).max() > 1e-6                  # ... and it is incomplete: fill-in positions are dropped


@pytest.mark.parametrize("system", SYSTEMS)
@pytest.mark.parametrize("pname,precond,sweeps", PRECONDS)
def test_first_iterate_is_the_preconditioner_application(system, pname, precond, sweeps):
    """After ONE CG iteration x_1 = lambda M^-1 b with lambda = (M^-1 b . b) / (M^-1 b . A M^-1 b): the oracle's preconditioner applied
    to b, seen through the solver interface, against the independent M^-1."""
    o, P, A, b = _system(system)
    x = np.zeros(P.ndof)
    o.solve(P.mat, b, x, method=capi.SOLVER_CG, precond=precond, sweeps=max(sweeps, 1), reduction=1e-30, maxit=1)
    q = _independent(A, pname, sweeps)(b)
    lam = (q @ b) / (q @ (A @ q))
    assert np.abs(x - lam * q).max() <= 1e-12 * np.abs(x).max()


@pytest.mark.parametrize("system", SYSTEMS)
@pytest.mark.parametrize("pname,precond,sweeps", PRECONDS)
def test_cg_iterates_and_counts(system, pname, precond, sweeps):
    """ISTL's CGSolver recurrences as the oracle states them against textbook PCG: the first iterates agree to rounding, the iteration
    counts to the given reduction agree (the systems are SPD up to the FD noise of the coupling entries and Dirichlet unit rows, H3)."""
    o, P, A, b = _system(system)
    M = _independent(A, pname, sweeps)
    xs, defects = pcg(A, b, M, 400, 1e-10)
    for k in (1, 2, 3, 5):
        if k > len(xs) - 1:                             # compare iterates before the textbook loop met the reduction
            break
        x = np.zeros(P.ndof)
        st = o.solve(P.mat, b, x, method=capi.SOLVER_CG, precond=precond, sweeps=max(sweeps, 1), reduction=1e-30, maxit=k)
        assert st.iterations == k
        assert np.abs(x - xs[k - 1]).max() <= 1e-10 * np.abs(xs[k - 1]).max(), k
        assert abs(st.defect - defects[k - 1]) <= 1e-9 * defects[0]
    x = np.zeros(P.ndof)
    st = o.solve(P.mat, b, x, method=capi.SOLVER_CG, precond=precond, sweeps=max(sweeps, 1), reduction=1e-10, maxit=400)
    assert st.converged and abs(st.iterations - len(xs)) <= 1, (st.iterations, len(xs))
    assert np.abs(x - sla.spsolve(A.tocsc(), b)).max() <= 1e-8 * np.abs(x).max()


@pytest.mark.parametrize("system", SYSTEMS)
@pytest.mark.parametrize("pname,precond,sweeps", PRECONDS)
def test_bicgstab_iterates_and_counts(system, pname, precond, sweeps):
    """ISTL's BiCGSTABSolver (right preconditioning, convergence test after both half steps) against van der Vorst's recurrences."""
    o, P, A, b = _system(system)
    M = _independent(A, pname, sweeps)
    xs, defects = bicgstab_right(A, b, M, 400, 1e-10)
    for k in (1, 2, 3):
        if k > len(xs) - 1:
            break
        x = np.zeros(P.ndof)
        st = o.solve(P.mat, b, x, method=capi.SOLVER_BICGSTAB, precond=precond, sweeps=max(sweeps, 1), reduction=1e-30, maxit=k)
        assert st.iterations == k
        assert np.abs(x - xs[k - 1]).max() <= 1e-9 * np.abs(xs[k - 1]).max(), k
    x = np.zeros(P.ndof)
    st = o.solve(P.mat, b, x, method=capi.SOLVER_BICGSTAB, precond=precond, sweeps=max(sweeps, 1), reduction=1e-10, maxit=400)
    assert st.converged and abs(st.iterations - len(xs)) <= 1, (st.iterations, len(xs))
    assert np.abs(x - sla.spsolve(A.tocsc(), b)).max() <= 1e-8 * np.abs(x).max()


def test_scipy_cg_takes_the_same_number_of_iterations():
    """A third opinion on the count: SciPy's own preconditioned CG with the independent Jacobi operator."""
    o, P, A, b = _system("testpoisson_3d")
    d = A.diagonal()
    count = [0]
    x, info = sla.cg(A, b, rtol=1e-10, atol=0.0, maxiter=400, M=sla.LinearOperator(A.shape, matvec=lambda r: r / d),
                     callback=lambda xk: count.__setitem__(0, count[0] + 1))
    assert info == 0
    z = np.zeros(P.ndof)
    st = o.solve(P.mat, b, z, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-10, maxit=400)
    assert st.converged and abs(st.iterations - count[0]) <= 1, (st.iterations, count[0])
    assert np.abs(z - x).max() <= 1e-8 * np.abs(x).max()
