"""Pins for the oracle's Neumann-Dirichlet coupling and Dirichlet-Neumann iteration — SURVEY §8 a10 / BASELINE configs[0]
(test/neumanndirichletcoupling.hh, test/testpoisson-dirichlet-neumann.cc:371-862 with conforming spaces).  No GPU."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import orc
import problems

capi = orc.capi


def test_operator_matrices_are_block_local_and_pattern_free():
    """LeftOperator = {left_sp, Coupling(left, right, ND)} (:566-578): the ND coupling declares no pattern
    (neumanndirichletcoupling.hh:51-52) and writes mat_ss only, so the matrix has entries in the left block only — which is
    what lets ISTL_SEQ_Subblock_Backend solve raw(A)[0][0] alone."""
    o = orc.Oracle()
    P = problems.dirichlet_neumann(o, (8, 6))
    off = o.get_child_offsets()
    for mat, k in ((P.lmat, 0), (P.rmat, 1)):
        rp, ci = o.matrix_get_pattern(mat)
        rows = np.repeat(np.arange(P.ndof), np.diff(rp))
        assert np.all((rows >= off[k]) & (rows < off[k + 1])) and np.all((ci >= off[k]) & (ci < off[k + 1]))
    # same pattern as the volume operator alone
    go_l = o.gridoperator_create([P.left], [])
    m = o.matrix_create([go_l])
    assert np.array_equal(o.matrix_get_pattern(m)[1], o.matrix_get_pattern(P.lmat)[1])


@pytest.mark.parametrize("n,fe", [((8, 6), (capi.FE_Q1, capi.FE_Q1)), ((4, 4), (capi.FE_Q2, capi.FE_Q1)), ((4, 3, 2), (capi.FE_Q1, capi.FE_Q1))])
@pytest.mark.parametrize("mode", [capi.ND_MODE_NEUMANN, capi.ND_MODE_DIRICHLET, capi.ND_MODE_BOTH])
def test_nd_jacobian_is_the_own_block_derivative(n, fe, mode):
    """jacobian_coupling (:291-432) = d r_s / d u_s exactly (the residual is affine); there is NO derivative with respect to the
    remote coefficients (mat_sr is a const reference): perturbing the right block changes the residual but not through J."""
    o = orc.Oracle()
    if len(n) == 3:
        pytest.importorskip("numpy")
    P = problems.dirichlet_neumann(o, n[:2], fe=fe, left_mode=capi.ND_MODE_DIRICHLET, right_mode=mode) if len(n) == 2 else None
    if P is None:                       # 3-D: same participants on a cube
        o.mesh_structured(n); o.subdomains_halfspace(0, 0.5, 0, 1)
        c0, c1 = o.add_space(fe[0], 0), o.add_space(fe[1], 1)
        pp = capi.PoissonParams(); pp.f = capi.Function(capi.FN_DN_SOURCE); pp.bc = capi.BCType(capi.BC_TESTPOISSON)
        pp.j = capi.Function(capi.FN_TESTPOISSON_J); pp.quad_order = 4
        l = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0]); r = o.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
        q = capi.NDCouplingParams(); q.mode, q.method, q.weights, q.intorderadd, q.alpha = mode, capi.DG_METHOD_SIPG, capi.DG_WEIGHTS_ON, 0, 2.0
        cp = o.add_coupling(capi.OP_ND_COUPLING, q, l, r)
        nd = o.finalize(); o.constraints_assemble([l, r], [pp.bc, pp.bc])
        from types import SimpleNamespace
        lop = o.gridoperator_create([l], [cp])
        P = SimpleNamespace(ndof=nd, lop=lop, lmat=o.matrix_create([lop]))
    off = o.get_child_offsets()
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, P.ndof)
    o.jacobian(P.lop, P.lmat, x, overwrite=True)
    A = o.csr(P.lmat)
    free = ~o.get_constraints().astype(bool)
    r0 = np.zeros(P.ndof); o.residual(P.lop, x, r0)
    dx = np.zeros(P.ndof); dx[:off[1]] = rng.uniform(-1, 1, off[1])
    r1 = np.zeros(P.ndof); o.residual(P.lop, x + dx, r1)
    assert np.abs((r1 - r0 - A @ dx)[free]).max() < 1e-12 * np.abs(r1 - r0).max()
    dy = np.zeros(P.ndof); dy[off[1]:off[2]] = rng.uniform(-1, 1, off[2] - off[1])
    r2 = np.zeros(P.ndof); o.residual(P.lop, x + dy, r2)
    assert np.abs(A @ dy).max() == 0.0 and np.abs(r2 - r0).max() > 1e-3


def test_penalty_block_closed_form_2d():
    """Dirichlet mode, alpha only: with theta-terms removed by symmetry check — the IP term of :429 is penalty * int phi_j phi_i over
    the face, penalty = (alpha / h_F) * 1 * k (k + d - 1), h_F = min(vol_s, vol_n) / vol_face = h (:128, :168).  Isolate it as
    J(alpha) - J(0)."""
    n = (8, 4)
    mats = []
    for alpha in (3.0, 0.0):
        o = orc.Oracle()
        P = problems.dirichlet_neumann(o, n, alpha=alpha, right_mode=capi.ND_MODE_DIRICHLET)
        o.jacobian(P.lop, P.lmat, np.zeros(P.ndof), overwrite=True)
        A = o.csr(P.lmat).toarray()
        con = o.get_constraints().astype(bool)
        mats.append((A, con))
    D = mats[0][0] - mats[1][0]
    con = mats[0][1]
    nx = n[0] // 2 + 1                          # left block lattice: nx x (n1 + 1), interface column ix = nx - 1
    hx, hy = 1.0 / n[0], 1.0 / n[1]
    pen = (3.0 / hx) * 1.0 * 1 * (1 + 2 - 1)    # h_F = hx*hy / hy
    ref = np.zeros_like(D)
    for e in range(n[1]):
        a, b = (nx - 1) + nx * e, (nx - 1) + nx * (e + 1)
        for (i, j, v) in ((a, a, 1 / 3), (b, b, 1 / 3), (a, b, 1 / 6), (b, a, 1 / 6)):
            ref[i, j] += pen * hy * v
    ref[con, :] = 0.0
    assert np.abs(D - ref).max() < 1e-12 * np.abs(ref).max()


def test_dn_iteration_converges_and_matches_direct_block_solves():
    """The loop of testpoisson-dirichlet-neumann.cc:721-822 with the shipped ini values converges (full-operator residual reduced by
    1e-8); with exact inner solves (reduction 1e-13) its iterates match the same loop run with SciPy direct block solves."""
    n = (16, 16)
    o = orc.Oracle()
    P = problems.dirichlet_neumann(o, n)
    u0 = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u0)
    hist = []
    u, its, r = problems.dn_iteration(o, P, u0.copy(), history=hist)
    assert its < 40 and hist[-1] / hist[0] < 1e-7
    res = np.zeros(P.ndof); o.residual(P.full, u, res)
    assert np.linalg.norm(res) == pytest.approx(r)

    tight = (1e-13, 1e-13, 0.2)
    h1, h2 = [], []
    u1, _, _ = problems.dn_iteration(o, P, u0.copy(), iterations=6, left_reduction=tight, right_reduction=tight, history=h1)
    off = o.get_child_offsets()

    def direct(mat, child, rr, z, reduction=None):
        A = o.csr(mat)[off[child]:off[child + 1], off[child]:off[child + 1]].tocsc()
        z[off[child]:off[child + 1]] = sla.spsolve(A, rr[off[child]:off[child + 1]])
    u2, _, _ = problems.dn_iteration(o, P, u0.copy(), iterations=6, solve_block=direct, history=h2)
    assert np.abs(u1 - u2).max() < 1e-9 * np.abs(u2).max()
    assert np.allclose(h1, h2, rtol=1e-6)
