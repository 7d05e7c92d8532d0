"""Pins for the CPU oracle (no GPU).  The reference ships no expected values ("parity unpinned"), so the
oracle is pinned by self-derived known answers: closed-form element matrices, an independent NumPy/SciPy construction of
the DOF numbering and sparsity pattern, algebraic identities of the assembled operators, FD-vs-exact
coupling Jacobians and SciPy direct solves."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import orc
import problems

capi = orc.capi


def test_q1_stiffness_2d_closed_form():
    o = orc.Oracle()
    P = problems.single_domain_poisson(o, (4, 4), quad_order=2)
    a = o.local_jacobian_volume(P.sp, 5, 4)
    ref = np.array([[4, -1, -1, -2], [-1, 4, -2, -1], [-1, -2, 4, -1], [-2, -1, -1, 4]]) / 6.0
    assert np.abs(a - ref).max() < 1e-15


def test_q1_stiffness_3d_closed_form():
    # trilinear Laplace element matrix on a cube of edge h: h * (1/3, 0, -1/12, -1/12) for (self, edge, face-diag, body-diag)
    o = orc.Oracle()
    n = 4
    P = problems.single_domain_poisson(o, (n, n, n), quad_order=4)
    a = o.local_jacobian_volume(P.sp, 21, 8)
    h = 1.0 / n
    for i in range(8):
        for j in range(8):
            dist = bin(i ^ j).count("1")
            ref = h * {0: 1 / 3.0, 1: 0.0, 2: -1 / 12.0, 3: -1 / 12.0}[dist]
            assert abs(a[i, j] - ref) < 1e-15


def test_q2_stiffness_rowsum_and_symmetry():
    o = orc.Oracle()
    P = problems.single_domain_poisson(o, (2, 2), quad_order=4, fe=capi.FE_Q2)
    a = o.local_jacobian_volume(P.sp, 0, 9)
    assert np.abs(a - a.T).max() < 1e-14
    assert np.abs(a.sum(axis=1)).max() < 1e-13          # gradients of a partition of unity sum to zero
    assert abs(np.trace(a) - (4 * 28 / 45 + 4 * 88 / 45 + 256 / 45)) < 1e-12   # biquadratic Laplace diagonal


def test_numbering_matches_appendix_b():
    """child block s = Q1 on subdomain s, global(s, v) = off_s + rank of v in V_s (ascending host
    vertex index), local dof i <-> corner with offset bits of i."""
    n = (6, 4, 5)
    o = orc.Oracle()
    P = problems.testpoisson(o, n, split_axis=1)
    ptr, dofs = o.get_dofmap()
    nx, ny, nz = n
    vid = lambda i, j, k: i + (nx + 1) * (j + (ny + 1) * k)
    in_sd = [set(), set()]
    cells = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                s = 1 if (j + 0.5) / ny > 0.5 else 0
                corners = [vid(i + (c & 1), j + ((c >> 1) & 1), k + ((c >> 2) & 1)) for c in range(8)]
                cells.append((s, corners))
                in_sd[s].update(corners)
    rank = [{v: r for r, v in enumerate(sorted(in_sd[s]))} for s in (0, 1)]
    off = [0, len(in_sd[0])]
    assert P.ndof == len(in_sd[0]) + len(in_sd[1])
    assert list(o.get_child_offsets()) == [0, len(in_sd[0]), P.ndof]
    for c, (s, corners) in enumerate(cells):
        assert ptr[c + 1] - ptr[c] == 8
        assert list(dofs[ptr[c]:ptr[c + 1]]) == [off[s] + rank[s][v] for v in corners]


def _independent_pattern(o, P, n, split_axis):
    ptr, dofs = o.get_dofmap()
    rows, cols = [], []
    dim = len(n)
    nn = list(n) + [1] * (3 - dim)
    nloc = 2 ** dim
    cell = lambda i, j, k: i + nn[0] * (j + nn[1] * k)
    for k in range(nn[2]):
        for j in range(nn[1]):
            for i in range(nn[0]):
                c = cell(i, j, k)
                d = dofs[ptr[c]:ptr[c + 1]]
                rows += list(np.repeat(d, nloc)); cols += list(np.tile(d, nloc))
                ijk = [i, j, k]
                if ijk[split_axis] == nn[split_axis] // 2 - 1:       # local (left) cell at the interface
                    m = list(ijk); m[split_axis] += 1
                    dn = dofs[ptr[cell(*m)]:ptr[cell(*m) + 1]]
                    rows += list(np.repeat(d, nloc)); cols += list(np.tile(dn, nloc))
                    rows += list(np.repeat(dn, nloc)); cols += list(np.tile(d, nloc))
    A = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(P.ndof, P.ndof)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    return A


@pytest.mark.parametrize("n,axis", [((8, 8), 1), ((6, 4), 0), ((4, 6, 4), 1), ((4, 4, 6), 2)])
def test_pattern_is_union_of_cliques(n, axis):
    o = orc.Oracle()
    P = problems.testpoisson(o, n, split_axis=axis)
    rowptr, colidx = o.matrix_get_pattern(P.mat)
    A = _independent_pattern(o, P, n, axis)
    assert np.array_equal(rowptr, A.indptr) and np.array_equal(colidx, A.indices)
    for r in range(P.ndof):
        assert np.all(np.diff(colidx[rowptr[r]:rowptr[r + 1]]) > 0)      # BCRS: ascending, no duplicates


def test_p2d_sizes_of_the_shipped_testpoisson():
    """testpoisson 4 2.0: 16x16 cells, Q1 on y<=.5 (17*9 dofs) and Q2 on y>=.5 (33*17 dofs): 714 in total."""
    o = orc.Oracle()
    P = problems.testpoisson(o, (16, 16), fe=(capi.FE_Q1, capi.FE_Q2))
    assert P.ndof == 714
    assert list(o.get_child_offsets()) == [0, 153, 714]
    assert P.ncon == 9 + 9 + 17          # x=0 on both subdomains, x=1 only where y <= .5


def test_jacobian_annihilates_constants_and_is_symmetric_on_free_rows():
    o = orc.Oracle()
    P = problems.testpoisson(o, (6, 6, 4), jac_mode=capi.JAC_ANALYTIC)
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    A = o.csr(P.mat)
    con = o.get_constraints().astype(bool)
    y = A @ np.ones(P.ndof)
    assert np.abs(y[~con]).max() < 1e-12 * np.abs(A.data).max()
    assert np.allclose(y[con], 1.0)
    F = A[~con][:, ~con]
    assert abs(F - F.T).max() < 1e-13 * np.abs(A.data).max()


def test_fd_coupling_jacobian_close_to_exact():
    oa, of = orc.Oracle(), orc.Oracle()
    Pa = problems.testpoisson(oa, (8, 8), jac_mode=capi.JAC_ANALYTIC)
    Pf = problems.testpoisson(of, (8, 8), jac_mode=capi.JAC_FD_BITMATCH)
    u = np.zeros(Pa.ndof); oa.interpolate(Pa.sps, Pa.gfn, u)
    oa.jacobian(Pa.go, Pa.mat, u, overwrite=True); of.jacobian(Pf.go, Pf.mat, u, overwrite=True)
    va, vf = oa.matrix_get_values(Pa.mat), of.matrix_get_values(Pf.mat)
    err = np.abs(va - vf).max() / np.abs(va).max()
    assert 0 < err < 1e-6            # FD noise is ~1e-8, and it is not identically the analytic matrix


def test_residual_is_linear_and_matches_jacobian():
    """All restated operators are linear: R(u) = J u + R(0) on unconstrained rows."""
    o = orc.Oracle()
    P = problems.testpoisson(o, (8, 6), jac_mode=capi.JAC_ANALYTIC)
    rng = np.random.default_rng(0)
    u = rng.uniform(-1, 1, P.ndof)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    A = o.csr(P.mat)
    r, r0 = np.zeros(P.ndof), np.zeros(P.ndof)
    o.residual(P.go, u, r); o.residual(P.go, np.zeros(P.ndof), r0)
    con = o.get_constraints().astype(bool)
    assert np.abs((A @ u + r0 - r)[~con]).max() < 1e-12 * np.abs(r).max()
    assert np.all(r[con] == 0.0)
    assert np.abs(r0).max() > 0        # the Neumann flux -5 on x=1, y>.5 is present


def test_patch_test_linear_solution():
    """A globally linear function has zero Laplace residual on every free row away from Neumann data and interface."""
    o = orc.Oracle()
    P = problems.single_domain_poisson(o, (5, 4, 3), quad_order=2)
    u = np.zeros(P.ndof)
    o.interpolate(P.sps, [capi.Function(capi.FN_POLY2, 1.0, 2.0, -3.0, 0.5, 0.0)], u)
    r = np.zeros(P.ndof); o.residual(P.go, u, r)
    assert np.abs(r).max() < 1e-13


def test_stationary_solve_against_scipy():
    """StationaryLinearProblemSolver semantics ([UPSTREAM] (vii)): r = R(u0), J z = r, u = u0 - z; BiCGSTAB + SSOR(3) as in
    testpoisson.cc:291-298, CG + Jacobi as the device backend does, SciPy's sparse LU as the second opinion."""
    o = orc.Oracle()
    P = problems.testpoisson(o, (16, 16))
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    r = np.zeros(P.ndof); o.residual(P.go, u, r)
    z_ref = sla.spsolve(o.csr(P.mat).tocsc(), r)
    z1, z2, z3 = np.zeros(P.ndof), np.zeros(P.ndof), np.zeros(P.ndof)
    s1 = o.solve(P.mat, r, z1, reduction=1e-13)
    s2 = o.solve(P.mat, r, z2, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-13)
    s3 = o.solve(P.mat, r, z3, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_ILU0, reduction=1e-13)
    for s, z in ((s1, z1), (s2, z2), (s3, z3)):
        assert s.converged
        assert np.abs(z - z_ref).max() <= 1e-10 * np.abs(z_ref).max()
    assert s1.iterations < s2.iterations


def test_mass_matrix_integrates_one():
    o = orc.Oracle()
    n = (4, 3, 2)
    o.mesh_structured(n)
    o.subdomains(np.ones(int(np.prod(n)), dtype=np.uint8))
    c0 = o.add_space(capi.FE_Q1, 0)
    lp = capi.L2Params(); lp.quad_order = 2
    s = o.add_subproblem(capi.OP_L2, lp, capi.COND_EQUALITY, 1, [c0])
    nd = o.finalize()
    go = o.gridoperator_create([s]); mat = o.matrix_create([go])
    o.jacobian(go, mat, np.zeros(nd), weight=2.5, overwrite=True)
    A = o.csr(mat)
    assert abs(A.sum() - 2.5) < 1e-13           # sum_ij weight * int phi_i phi_j = weight * |Omega|


def test_parallel_baseline_assembly_equals_serial():
    o1, o2 = orc.Oracle(), orc.Oracle()
    P1, P2 = problems.testpoisson(o1, (6, 6, 12)), problems.testpoisson(o2, (6, 6, 12))
    u = np.zeros(P1.ndof); o1.interpolate(P1.sps, P1.gfn, u)
    o1.jacobian(P1.go, P1.mat, u, overwrite=True)
    o2.jacobian(P2.go, P2.mat, u, overwrite=True, nthreads=4)
    a, b = o1.matrix_get_values(P1.mat), o2.matrix_get_values(P2.mat)
    assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()


def test_implicit_euler_against_scipy():
    """The two-operator driver of tests/problems.implicit_euler_step (OneStepGridOperator restated) against a direct SciPy
    solve of (M + dt K) u_new = M u_old + dt b assembled from the oracle's matrices: R is affine in u."""
    o = orc.Oracle()
    P = problems.instationary(o, (10, 8))
    dt, t1 = 0.05, 0.05
    u0 = np.zeros(P.ndof)
    o.interpolate(P.sps, P.gfn, u0, time=0.0)
    u1 = problems.implicit_euler_step(o, P, u0, t1, dt, lambda mat, R, z: o.solve(mat, R, z, reduction=1e-13))
    # independent: J = J1 + dt J0 (already in the matrix), R(u) evaluated at the Dirichlet-updated guess
    J = o.csr(P.mat)
    guess = np.zeros(P.ndof)
    o.interpolate(P.sps, P.gfn, guess, time=t1)
    o.copy_nonconstrained(u0, guess)
    R = np.zeros(P.ndof)
    o.residual(P.go1, guess, R, weight=1.0)
    o.residual(P.go1, u0, R, weight=-1.0)
    o.residual(P.go0, guess, R, weight=dt)
    z = sla.spsolve(sp.csc_matrix(J), R)
    assert np.abs((guess - z) - u1).max() <= 1e-10 * np.abs(u1).max()
    con = o.get_constraints().astype(bool)
    assert np.array_equal(u1[con], guess[con])            # Dirichlet values come from the interpolation at the new time
    # the new iterate satisfies the discrete equation on free rows: R(u1) = 0
    R1 = np.zeros(P.ndof)
    o.residual(P.go1, u1, R1, weight=1.0)
    o.residual(P.go1, u0, R1, weight=-1.0)
    o.residual(P.go0, u1, R1, weight=dt)
    assert np.abs(R1).max() <= 1e-9 * np.abs(R).max()


def test_one_step_schemes_have_their_orders():
    """[UPSTREAM] ImplicitEulerParameter, OneStepThetaParameter(1/2) and Alexander2Parameter (alpha = 1 - sqrt(2)/2) as
    tests/problems.one_step_method restates them (the stage equations the device's mdb_onestep_apply is checked against): on the
    instationary two-subdomain problem with time-dependent Dirichlet data the error against a fine-step solution falls with order 1,
    2 and 2.  Wrong stage coefficients, stage times or weights of the two grid operators lose the order."""
    T = 0.4

    def run(scheme, theta, nsteps):
        o = orc.Oracle()
        P = problems.instationary(o, (10, 8))
        u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u, time=0.0)
        dt = T / nsteps
        for s in range(nsteps):
            u = problems.one_step_method(o, P, u, s * dt, dt, lambda mat, R, z: o.solve(mat, R, z, reduction=1e-14), scheme, theta)
        return u

    ref = run("alexander2", 0.0, 1024)
    for scheme, theta, steps, order in (("implicit_euler", 1.0, (16, 32, 64), 1.0), ("theta", 0.5, (64, 128, 256), 2.0),
                                        ("alexander2", 0.0, (8, 16, 32), 2.0)):
        e = [np.abs(run(scheme, theta, n) - ref).max() for n in steps]
        rates = [np.log2(e[i] / e[i + 1]) for i in range(2)]
        assert all(abs(r - order) < 0.3 for r in rates), (scheme, e, rates)
    # theta = 1 is implicit Euler, and both agree with the dedicated two-operator driver of implicit_euler_step
    a, b = run("implicit_euler", 1.0, 4), run("theta", 1.0, 4)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


@pytest.mark.parametrize("n,axis", [((6, 8), 1), ((8, 6), 0), ((4, 3, 6), 2), ((3, 6, 4), 1)])
def test_volume_blocks_are_kronecker_products(n, axis):
    """Numbering and assembly pinned together by a construction that shares nothing with a cell loop: on a uniform mesh the Q1 Poisson
    stiffness of a box is a sum of Kronecker products of the 1-D stiffness and mass matrices, K = sum_d (x)_e (e == d ? K1_e : M1_e),
    indexed lexicographically with axis 0 fastest.  Each child block of the oracle's Jacobian (rows and columns of one subdomain's space:
    SURVEY Appendix B — child blocks concatenated, each numbered lexicographically over its subdomain's vertices) must equal it on every
    row that is neither Dirichlet-constrained (unit rows) nor within one cell layer of the interface (the coupling's ss / nn blocks)."""
    o = orc.Oracle()
    P = problems.testpoisson(o, n, split_axis=axis)
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    A = o.csr(P.mat)
    con = o.get_constraints().astype(bool)
    off = o.get_child_offsets()
    dim = len(n)

    def k1(m, h):
        return sp.diags([-np.ones(m), np.r_[1.0, 2.0 * np.ones(m - 1), 1.0], -np.ones(m)], [-1, 0, 1]) / h

    def m1(m, h):
        return sp.diags([np.ones(m), np.r_[2.0, 4.0 * np.ones(m - 1), 2.0], np.ones(m)], [-1, 0, 1]) * (h / 6.0)

    checked = 0
    for child in (0, 1):
        cells = list(n)
        cells[axis] = n[axis] // 2                                  # each subdomain is half the box along the split axis
        h = [1.0 / n[d] for d in range(dim)]
        K = None
        for d in range(dim):
            term = None
            for e in range(dim):                                    # axis 0 fastest: kron(last axis, ..., axis 0)
                f = (k1 if e == d else m1)(cells[e], h[e])
                term = f if term is None else sp.kron(f, term)
            K = term if K is None else K + term
        K = sp.csr_matrix(K)
        size = int(np.prod([c + 1 for c in cells]))
        assert off[child + 1] - off[child] == size
        B = A[off[child]:off[child + 1], off[child]:off[child + 1]].toarray()
        # lattice coordinate along the split axis of every DOF of the block
        idx = np.arange(size)
        stride = int(np.prod([c + 1 for c in cells[:axis]]))
        j = (idx // stride) % (cells[axis] + 1)
        near = (j >= cells[axis] - 1) if child == 0 else (j <= 1)   # the two vertex planes of the cell layer at the interface
        rows = ~near & ~con[off[child]:off[child + 1]]
        assert rows.sum() > 0
        scale = np.abs(K.data).max()
        assert np.abs(B[rows] - K.toarray()[rows]).max() <= 1e-13 * scale
        # and those rows hold nothing outside their own block
        full = A[off[child]:off[child + 1]].toarray()
        full[:, off[child]:off[child + 1]] = 0.0
        assert np.abs(full[rows]).max() == 0.0
        checked += int(rows.sum())
    assert checked > 0


def test_mass_blocks_are_kronecker_products():
    """The L2 operator of the instationary driver (dt1 grid operator, test/testinstationarypoisson.cc:225-226) the same way: every child
    block of its matrix is the Kronecker product of the 1-D Q1 mass matrices — on ALL rows (no coupling, no Dirichlet handling enters
    the mass operator's volume term; constrained rows are replaced by unit rows, so they are left out)."""
    o = orc.Oracle()
    n = (6, 4)
    P = problems.instationary(o, n, split_axis=0)
    u = np.zeros(P.ndof)
    o.jacobian(P.go1, P.mat, u, overwrite=True)
    A = o.csr(P.mat)
    con = o.get_constraints().astype(bool)
    off = o.get_child_offsets()

    def m1(m, h):
        return sp.diags([np.ones(m), np.r_[2.0, 4.0 * np.ones(m - 1), 2.0], np.ones(m)], [-1, 0, 1]) * (h / 6.0)
    for child in (0, 1):
        cells = [n[0] // 2, n[1]]
        M = sp.kron(m1(cells[1], 1.0 / n[1]), m1(cells[0], 1.0 / n[0])).toarray()
        B = A[off[child]:off[child + 1], off[child]:off[child + 1]].toarray()
        rows = ~con[off[child]:off[child + 1]]
        assert rows.sum() > 0 and np.abs(B[rows] - M[rows]).max() <= 1e-14 * np.abs(M).max()


@pytest.mark.parametrize("n,axis,intensity", [((6, 4), 1, 2.0), ((4, 8), 0, 3.5), ((4, 2, 4), 2, 2.0), ((2, 4, 4), 0, 1.25)])
def test_cvcf_penalty_and_flux_terms_in_closed_form(n, axis, intensity):
    """ContinuousValueContinuousFlowCoupling (test/proportionalflowcoupling.hh:113-233) on the jump u = 0 | 1 across the interface, all
    boundaries Neumann, no source: the volume terms vanish, the flux average vanishes (grad u = 0 on both sides), and what is left is
    r1_i = int_Gamma [ -1/2 (grad phi1_i . n1) [u] + (alpha / h_F) [u] phi1_i ],  [u] = u1 - u2 = -1  (mirrored on side 2).
    Because the shape functions of a side sum to one, the rows of a side add up to  -+ (alpha / h_F) |Gamma|  exactly — this pins the
    penalty coefficient with h_F = |corner 0 - corner 1| of the face (:135), which the patch and symmetry tests cannot see (the penalty
    vanishes on continuous functions).  In 2-D the single rows are closed forms too."""
    o = orc.Oracle()
    dim = len(n)
    o.mesh_structured(n)
    o.subdomains_halfspace(axis, 0.5, 0, 1)
    c0, c1 = o.add_space(capi.FE_Q1, 0), o.add_space(capi.FE_Q1, 1)
    pp = capi.PoissonParams()
    pp.f = capi.Function(capi.FN_ZERO); pp.bc = capi.BCType(capi.BC_ALL_NEUMANN); pp.j = capi.Function(capi.FN_ZERO); pp.quad_order = 4
    left = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0])
    right = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
    cv = capi.CVCFParams(); cv.quad_order = 4; cv.intensity = intensity
    cp = o.add_coupling(capi.OP_CVCF_COUPLING, cv, left, right, capi.JAC_FD_BITMATCH)
    ndof = o.finalize()
    assert o.constraints_assemble([left, right], [pp.bc, pp.bc]) == 0
    go = o.gridoperator_create([left, right], [cp])
    off = o.get_child_offsets()
    u = np.zeros(ndof); u[off[1]:off[2]] = 1.0
    r = np.zeros(ndof); o.residual(go, u, r)
    h = [1.0 / m for m in n]
    # the face's corner 0 -> corner 1 runs along the first axis of the face, i.e. the lowest axis that is not the normal
    first_face_axis = 0 if axis != 0 else 1
    h_F = h[first_face_axis]
    gamma = 1.0                                             # area of the interface of the unit box
    want = intensity / h_F * gamma
    assert abs(r[off[0]:off[1]].sum() + want) <= 1e-12 * want
    assert abs(r[off[1]:off[2]].sum() - want) <= 1e-12 * want
    if dim == 2:
        # single rows of side 1 (child 0): interface vertices carry -alpha / h_F * int phi + 1/2 int d_n phi, the layer behind them the
        # gradient term alone with the opposite sign; int phi over the faces at an interior interface vertex = h_t, at an end vertex h_t / 2
        t = 1 - axis
        ht, hn = h[t], h[axis]
        cells = list(n); cells[axis] = n[axis] // 2
        ext = [c + 1 for c in cells]
        blk = r[off[0]:off[1]].reshape(ext[1], ext[0])      # axis 0 fastest
        top = blk[-1, :] if axis == 1 else blk[:, -1]
        behind = blk[-2, :] if axis == 1 else blk[:, -2]
        w = np.full(ext[t], ht); w[0] = w[-1] = ht / 2.0
        assert np.abs(top - (-(intensity / h_F) * w + 0.5 * w / hn)).max() <= 1e-12 * want
        assert np.abs(behind - (-0.5 * w / hn)).max() <= 1e-12 * want
        rest = (blk[:-2, :] if axis == 1 else blk[:, :-2])
        assert np.abs(rest).max() <= 1e-13 * want


@pytest.mark.parametrize("n,axis", [((6, 4), 0), ((2, 4, 4), 2)])
def test_proportional_flow_coupling_in_closed_form(n, axis):
    """ProportionalFlowCoupling (test/proportionalflowcoupling.hh:27-86): r_s_i += k (u_s - u_n) phi_s_i, r_n_i -= k (u_s - u_n) phi_n_i on the
    interface.  For the jump u = 0 | 1 the rows of the interface vertices are -+ k int phi_i (the trace mass of the vertex) and every other
    row vanishes; its Jacobian blocks are -+ k times the trace mass matrix (checked through J u = r: the residual is linear)."""
    o = orc.Oracle()
    k = 2.5
    o.mesh_structured(n)
    o.subdomains_halfspace(axis, 0.5, 0, 1)
    c0, c1 = o.add_space(capi.FE_Q1, 0), o.add_space(capi.FE_Q1, 1)
    pp = capi.PoissonParams()
    pp.f = capi.Function(capi.FN_ZERO); pp.bc = capi.BCType(capi.BC_ALL_NEUMANN); pp.j = capi.Function(capi.FN_ZERO); pp.quad_order = 4
    left = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0])
    right = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
    pf = capi.PropFlowParams(); pf.intensity = k
    cp = o.add_coupling(capi.OP_PROPFLOW_COUPLING, pf, left, right, capi.JAC_FD_BITMATCH)
    ndof = o.finalize()
    o.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = o.gridoperator_create([left, right], [cp])
    mat = o.matrix_create([go])
    off = o.get_child_offsets()
    u = np.zeros(ndof); u[off[1]:off[2]] = 1.0
    r = np.zeros(ndof); o.residual(go, u, r)
    dim = len(n)
    h = [1.0 / m for m in n]
    cells = list(n); cells[axis] = n[axis] // 2
    ext = [c + 1 for c in cells]
    # trace mass of a vertex: product over the tangential axes of h (interior) or h / 2 (end of the interface)
    w = None
    for d in reversed(range(dim)):                           # indexed [highest tangential axis, ..., lowest] like the reshape below
        if d == axis:
            continue
        wd = np.full(ext[d], h[d]); wd[0] = wd[-1] = h[d] / 2.0
        w = wd if w is None else np.multiply.outer(w, wd)
    shape = tuple(reversed(ext))
    b0 = r[off[0]:off[1]].reshape(shape)
    b1 = r[off[1]:off[2]].reshape(shape)
    ax = dim - 1 - axis                                      # position of the split axis in the reshaped array
    top0 = np.take(b0, ext[axis] - 1, axis=ax)               # subdomain 0's interface plane is its last one, subdomain 1's its first
    bot1 = np.take(b1, 0, axis=ax)
    assert w.shape == top0.shape
    assert np.abs(top0 + k * w).max() <= 1e-13 * k and np.abs(bot1 - k * w).max() <= 1e-13 * k
    assert abs(np.abs(r).sum() - 2.0 * k) <= 1e-12 * k       # nothing else: the two planes carry k |Gamma| each
    o.jacobian(go, mat, u, overwrite=True)
    assert np.abs(o.csr(mat) @ u - r).max() <= 1e-8 * k      # FD Jacobian of a linear residual (epsilon = 1e-8)


@pytest.mark.parametrize("n,axis", [((4, 6), 1), ((6, 4), 0), ((2, 3, 4), 2)])
def test_source_and_neumann_terms_in_closed_form(n, axis):
    """Poisson's lambda terms ([UPSTREAM] PDELab::Poisson; in-tree twin test/instationarypoissonoperator.hh:103-117, :142-167) with a
    constant source c and a constant Neumann flux q on the whole outer boundary, at u = 0:  r_i = -c int phi_i + q int_{dOmega} phi_i.
    On a uniform mesh both integrals are products of 1-D trapezoid weights over the SUBDOMAIN's lattice.  The interface is an interior
    face on which the assembler calls the subproblem's boundary terms (residualassemblerfunctors.hh:80-86) and the boundary-type function
    answers "none" (test/testpoisson.cc:54-57): a Neumann term there would add q |Gamma| to the interface rows — the closed form says no."""
    c, q = 1.75, -0.6
    o = orc.Oracle()
    dim = len(n)
    o.mesh_structured(n)
    o.subdomains_halfspace(axis, 0.5, 0, 1)
    c0, c1 = o.add_space(capi.FE_Q1, 0), o.add_space(capi.FE_Q1, 1)
    pp = capi.PoissonParams()
    pp.f = capi.Function(capi.FN_CONST, c); pp.bc = capi.BCType(capi.BC_ALL_NEUMANN); pp.j = capi.Function(capi.FN_CONST, q); pp.quad_order = 4
    left = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0])
    right = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
    cv = capi.CVCFParams(); cv.quad_order = 4; cv.intensity = 2.0
    cp = o.add_coupling(capi.OP_CVCF_COUPLING, cv, left, right, capi.JAC_FD_BITMATCH)
    ndof = o.finalize()
    o.constraints_assemble([left, right], [pp.bc, pp.bc])
    go = o.gridoperator_create([left, right], [cp])
    off = o.get_child_offsets()
    r = np.zeros(ndof); o.residual(go, np.zeros(ndof), r)
    h = [1.0 / m for m in n]
    total = 0.0
    for child in (0, 1):
        cells = list(n); cells[axis] = n[axis] // 2
        ext = [m + 1 for m in cells]
        w = []
        for d in range(dim):
            wd = np.full(ext[d], h[d]); wd[0] = wd[-1] = h[d] / 2.0
            w.append(wd)
        # is the lattice plane an OUTER boundary of the unit box?  (the interface plane of the split axis is not)
        outer = []
        for d in range(dim):
            lo, hi = np.zeros(ext[d], dtype=bool), np.zeros(ext[d], dtype=bool)
            lo[0] = not (d == axis and child == 1)
            hi[-1] = not (d == axis and child == 0)
            outer.append(lo | hi)

        def tensor(factors):                             # array indexed [axis dim-1, ..., axis 0]
            t = None
            for d in reversed(range(dim)):
                t = factors[d] if t is None else np.multiply.outer(t, factors[d])
            return t
        W = tensor(w)
        B = np.zeros_like(W)
        for d in range(dim):
            B += tensor([outer[e].astype(float) if e == d else w[e] for e in range(dim)])
        want = -c * W + q * B
        got = r[off[child]:off[child + 1]].reshape(W.shape)
        assert np.abs(got - want).max() <= 1e-13 * max(abs(c), abs(q)), child
        total += got.sum()
    area = 2.0 * dim                                     # surface of the unit box
    assert abs(total - (-c * 1.0 + q * area)) <= 1e-12


@pytest.mark.parametrize("n", [(8, 8), (4, 6), (4, 4, 6)])
def test_dirichlet_set_unit_rows_and_interpolated_values(n):
    """Constraints assembly (constraints.hh:312-558, constraintsfunctors.hh:41-131) for the testpoisson boundary types (test/testpoisson.cc:
    49-71, evaluated at the FACE CENTRE; a Dirichlet face constrains all its vertices): the exact set from the lattice — x_0 = 0 everywhere,
    x_0 = 1 where the face centre has x_1 < 0.5, nothing on x_1 = 0 / 1 and (3-D) nothing on the x_2 faces beyond what those faces' own
    centres give — then [UPSTREAM] handle_dirichlet_constraints: constrained rows are unit rows while the COLUMNS of constrained DOFs stay
    (the matrix is not symmetrised), and interpolation puts G(vertex) into the constrained coefficients."""
    o = orc.Oracle()
    dim = len(n)
    P = problems.testpoisson(o, n, split_axis=1)
    con = o.get_constraints().astype(bool)
    off = o.get_child_offsets()
    want = np.zeros(P.ndof, dtype=bool)
    coords = np.zeros((P.ndof, dim))
    for child in (0, 1):
        cells = list(n); cells[1] = n[1] // 2
        ext = [m + 1 for m in cells]
        idx = np.arange(int(np.prod(ext)))
        ijk = []
        rem = idx.copy()
        for d in range(dim):
            ijk.append(rem % ext[d]); rem //= ext[d]
        gj = ijk[1] + (n[1] // 2 if child == 1 else 0)                 # global vertex index along the split axis
        x0_lo = ijk[0] == 0
        # a coefficient on x_0 = 1 is constrained iff one of the faces there OF ITS OWN SUBDOMAIN'S CELLS has its centre below x_1 = 0.5:
        # every such coefficient of the lower subdomain (its copy of the interface vertex included), none of the upper one
        x0_hi = (ijk[0] == n[0]) & (child == 0)
        flag = x0_lo | x0_hi
        if dim == 3:
            # faces x_2 = 0 / 1: centre x_1 in (0, 1) is never within 1e-6 of the ends, and the x_0 > 1 - 1e-6 test fails at a face centre
            # (x_0 centre < 1): Dirichlet.  All their vertices are constrained.
            flag |= (ijk[2] == 0) | (ijk[2] == n[2])
        want[off[child]:off[child + 1]] = flag
        g = [ijk[0] / n[0], gj / n[1]] + ([ijk[2] / n[2]] if dim == 3 else [])
        coords[off[child]:off[child + 1]] = np.stack(g, axis=1)
    assert np.array_equal(con, want)
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    G = np.exp(-((coords - 0.5) ** 2).sum(axis=1))                      # G: exp(-|x - 0.5|^2), testpoisson.cc:92-99
    assert np.abs(u - G).max() <= 1e-15
    o.jacobian(P.go, P.mat, u, overwrite=True)
    A = o.csr(P.mat).toarray()
    assert np.array_equal(A[con], np.eye(P.ndof)[con])                  # unit rows
    assert np.abs(A[~con][:, con]).max() > 0.01                         # the columns of constrained DOFs are kept (no symmetrisation)
