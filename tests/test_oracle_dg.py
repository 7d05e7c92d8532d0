"""Pins for the oracle's DG path — BASELINE configs[0] as shipped (test/testpoisson-dirichlet-neumann.cc:883-977: QkDG spaces,
[UPSTREAM] ConvectionDiffusionDG(SIPG, weightsOn, alpha = 2) on both subdomains, ContinuousValueContinuousFlowCoupling between them)
and the alpha_skeleton / alpha_boundary-on-interface dispatch of residualassemblerfunctors.hh:58-120.  No GPU.  The reference ships no
expected values, so the restatement is pinned by independent constructions: numbering and pattern from the cell adjacency, closed
forms of the penalty blocks, symmetry of SIPG, consistency (the exact polynomial solution has zero residual), affine residual,
SciPy solves and the convergence order of the discretisation."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import orc
import problems

capi = orc.capi


def _setup(n, **kw):
    o = orc.Oracle()
    return o, problems.dirichlet_neumann_dg(o, n, **kw)


def test_sizes_of_the_shipped_configuration():
    """ini mesh.refine = 9: 2 * (512 * 256) cells * 9 = 2 359 296 DOFs; checked here at refine 4 (same formula), no constrained DOF."""
    o, P = _setup((16, 16))
    assert P.ndof == 2 * 16 * 8 * 9 and P.ncon == 0
    assert list(o.get_child_offsets()) == [0, 16 * 8 * 9, 2 * 16 * 8 * 9]
    assert 2 * 512 * 256 * 9 == 2359296


@pytest.mark.parametrize("n,k", [((6, 4), 2), ((4, 5), 1), ((4, 3, 2), 1), ((2, 2, 3), 2)])
def test_dg_numbering_and_pattern_from_cell_adjacency(n, k):
    """DOF = offset(subdomain) + (rank of the cell among the subdomain's cells) * (k+1)^d + i; pattern = cell block (FullVolumePattern)
    + blocks of the face neighbours in the same subdomain (FullSkeletonPattern) + blocks across the interface (FullCouplingPattern):
    every row of a cell with m face neighbours has (m + 1) (k+1)^d entries — 45 for an interior DG-Q2 cell in 2-D (ini:5-7 hint)."""
    o, P = _setup(n, k=k)
    dim = len(n); nl = (k + 1) ** dim
    ptr, dofs = o.get_dofmap()
    ncell = int(np.prod(n))
    idx = np.indices(n[::-1])[::-1].reshape(dim, -1)          # idx[d][c]
    right = idx[0] >= n[0] // 2
    rank = np.where(right, np.cumsum(right) - 1, np.cumsum(~right) - 1)
    off = o.get_child_offsets()
    first = np.where(right, off[1], off[0]) + rank * nl
    assert np.all(np.diff(ptr) == nl)
    assert np.array_equal(dofs.reshape(ncell, nl), first[:, None] + np.arange(nl)[None, :])
    rows, cols = [], []
    strides = np.cumprod([1] + list(n[:-1]))
    for c in range(ncell):
        nb = [c]
        for d in range(dim):
            if idx[d][c] > 0: nb.append(c - strides[d])
            if idx[d][c] < n[d] - 1: nb.append(c + strides[d])
        for b in nb:
            rows.append(np.repeat(first[c] + np.arange(nl), nl)); cols.append(np.tile(first[b] + np.arange(nl), nl))
    ref = sp.csr_matrix((np.ones(sum(len(r) for r in rows)), (np.concatenate(rows), np.concatenate(cols))), shape=(P.ndof, P.ndof))
    ref.sum_duplicates(); ref.sort_indices()
    rp, ci = o.matrix_get_pattern(P.mat)
    assert np.array_equal(rp, ref.indptr) and np.array_equal(ci, ref.indices)
    if n == (6, 4):
        assert np.diff(rp).max() == 45


def test_penalty_blocks_closed_form():
    """penalty_factor = alpha / h_F * harmonic_average * k (k + d - 1) with h_F = min(vol_s, vol_n) / vol_face: on a 2 x 2 grid (h = 1/2,
    Q1, all-Neumann boundary so that only the two skeleton faces carry a penalty) J(alpha = 3) - J(alpha = 1) is 2 * (2 / h) times
    the jump mass matrix of the faces y = 1/2 inside each subdomain: +-(h/3, h/6) on the four DOFs of the face."""
    n, h = (2, 2), 0.5
    mats = []
    for alpha in (1.0, 3.0):
        o, P = _setup(n, k=1, alpha=alpha, coupled=False, bc=capi.BCType(capi.BC_ALL_NEUMANN))
        o.jacobian(P.go, P.mat, np.zeros(P.ndof), overwrite=True)
        mats.append(o.csr(P.mat).toarray())
    D = mats[1] - mats[0]
    ref = np.zeros_like(D)
    M = h * np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    for sd in (0, 1):
        lo, up = sd * 8 + 0, sd * 8 + 4                     # the subdomain's lower cell (rank 0) and upper cell (rank 1)
        ds = np.array([lo + 2, lo + 3])                     # lower cell: local DOFs 2, 3 on its top face
        dn = np.array([up + 0, up + 1])                     # upper cell: local DOFs 0, 1 on its bottom face
        fac = 2.0 * (1.0 / h) * 1 * (1 + 2 - 1)
        ref[np.ix_(ds, ds)] += fac * M; ref[np.ix_(dn, dn)] += fac * M
        ref[np.ix_(ds, dn)] -= fac * M; ref[np.ix_(dn, ds)] -= fac * M
    assert np.abs(D - ref).max() < 1e-13


@pytest.mark.parametrize("n,k", [((6, 4), 1), ((4, 4), 2), ((3, 2, 2), 2)])
def test_sipg_is_symmetric_and_nipg_is_not(n, k):
    o, P = _setup(n, k=k, coupled=False)
    o.jacobian(P.go, P.mat, np.zeros(P.ndof), overwrite=True)
    A = o.csr(P.mat)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max()
    o2, P2 = _setup(n, k=k, coupled=False, method=capi.DG_METHOD_NIPG)
    o2.jacobian(P2.go, P2.mat, np.zeros(P2.ndof), overwrite=True)
    B = o2.csr(P2.mat)
    assert abs(B - B.T).max() > 1e-3 * abs(B).max()
    # NIPG and SIPG differ only in the sign of the consistency-transposed term: their mean is the part common to both
    assert abs((A + B) / 2 - (A + B).T / 2).max() > 0


@pytest.mark.parametrize("n,k,coupled", [((4, 4), 1, True), ((4, 6), 2, True), ((4, 3), 2, False), ((2, 3, 2), 2, True), ((4, 2, 2), 1, True)])
def test_consistency_exact_polynomial_has_zero_residual(n, k, coupled):
    """u in the discrete space with -Laplace u = f, Dirichlet data g = u: every jump vanishes and the flux terms integrate the volume
    term by parts, so R(u) = 0 — on the skeleton faces (alpha_skeleton), on the boundary (alpha_boundary, Dirichlet branch), and on
    the subdomain interface, where the subproblems' alpha_boundary must do nothing (bctype None) and the coupling operator supplies
    the flux.  Uncoupled, the interface has no flux term at all and the residual is NOT zero there."""
    dim = len(n)
    if k == 1:
        u = capi.Function(capi.FN_POLY2, 0.3, 1.0, -2.0, 0.5, 0.0); f = capi.Function(capi.FN_ZERO)
    else:
        u = capi.Function(capi.FN_POLY2, 0.3, 1.0, -2.0, 0.5, 1.5); f = capi.Function(capi.FN_CONST, -2.0 * dim * 1.5)
    o, P = _setup(n, k=k, f=f, g=u, bc=capi.BCType(capi.BC_ALL_DIRICHLET), coupled=coupled)
    x = np.zeros(P.ndof); o.interpolate(P.sps, [u, u], x)
    r = np.zeros(P.ndof); o.residual(P.go, x, r)
    if coupled:
        assert np.abs(r).max() < 1e-12
    else:
        assert np.abs(r).max() > 1e-3


@pytest.mark.parametrize("n,k", [((6, 4), 2), ((3, 3, 2), 1)])
def test_residual_is_affine_and_fd_coupling_jacobian(n, k):
    o, P = _setup(n, k=k)
    x = np.random.default_rng(0).uniform(-1, 1, P.ndof)
    o.jacobian(P.go, P.mat, x, overwrite=True)
    A = o.csr(P.mat)
    r0, rx = np.zeros(P.ndof), np.zeros(P.ndof)
    o.residual(P.go, np.zeros(P.ndof), r0); o.residual(P.go, x, rx)
    assert np.abs(A @ x + r0 - rx).max() < 5e-6 * np.abs(rx).max()          # the coupling blocks are FD (eps = 1e-8)


def test_solve_against_scipy_and_convergence_order():
    """BiCGSTAB + SSOR(3) (ISTLBackend_SEQ_BCGS_SSOR) against spsolve; DG-Q2 on a smooth manufactured solution (u = Gauss bump as
    Dirichlet data, f = -Laplace u is not in the closed function set, so: u = POLY2, exact) and, for the order, Q1 against Q2 errors."""
    errs = {}
    for k in (1, 2):
        for m in (4, 8):
            n = (m, m)
            u = capi.Function(capi.FN_POLY2, 0.0, 0.0, 0.0, 0.0, 1.0)       # |x|^2: in Q2, not in Q1
            f = capi.Function(capi.FN_CONST, -4.0)
            o, P = _setup(n, k=k, f=f, g=u, bc=capi.BCType(capi.BC_ALL_DIRICHLET))
            x0 = np.zeros(P.ndof)
            r = np.zeros(P.ndof); o.residual(P.go, x0, r)
            o.jacobian(P.go, P.mat, x0, overwrite=True)
            z = np.zeros(P.ndof)
            st = o.solve(P.mat, r, z, reduction=1e-12)
            assert st.converged
            zs = sla.spsolve(o.csr(P.mat).tocsc(), r)
            assert np.abs(z - zs).max() < 1e-8 * np.abs(zs).max()
            ex = np.zeros(P.ndof); o.interpolate(P.sps, [u, u], ex)
            errs[k, m] = np.abs((x0 - zs) - ex).max()
    assert errs[2, 4] < 1e-10 and errs[2, 8] < 1e-10                        # exact in Q2
    assert errs[1, 8] < 0.4 * errs[1, 4]                                    # second order in Q1 (nodal error)


def test_dirichlet_neumann_iteration_on_the_shipped_dg_spaces():
    """solve() as main() instantiates it (DG-Q2 + ConvectionDiffusionDG, :968-971) with the Neumann-Dirichlet couplings: the loop
    of :721-822 reaches the 1e-8 reduction of the full-operator residual, and its limit and the monolithic solution (a different,
    symmetric coupling operator) are two discretisations of the same problem: they approach each other under refinement."""
    diffs = []
    for m in (8, 16):
        o = orc.Oracle()
        P = problems.dirichlet_neumann(o, (m, m), fe=(capi.FE_DGQ2, capi.FE_DGQ2), dg=True)
        u0 = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u0)
        h = []
        u, its, _ = problems.dn_iteration(o, P, u0.copy(), history=h)
        assert its < 40 and h[-1] <= 1e-8 * 40 * h[0]
        o.jacobian(P.go, P.mat, u0, overwrite=True)
        r = np.zeros(P.ndof); o.residual(P.go, u0, r)
        z = sla.spsolve(o.csr(P.mat).tocsc(), r)
        diffs.append(np.abs((u0 - z) - u).max())
    assert diffs[1] < 0.7 * diffs[0]
