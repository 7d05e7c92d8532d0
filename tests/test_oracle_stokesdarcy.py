"""Pins of oracle/um_stokesdarcy.py (groundwork for SURVEY §8 row a12: the reference's StokesDarcyCouplingOperator,
test/stokesdarcycouplingoperator.hh:104-245, restated for one interface edge; no device path yet).  Known answers only — the
reference's tests hold none: identities of the Pk2D bases, closed-form row sums of the three coupling terms, and the hydrostatic
state, which is an equilibrium of the intended operator but NOT of the operator as written (the vsize / dsize loop quirk)."""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import um_stokesdarcy as sd  # noqa: E402


@pytest.mark.parametrize("k", [1, 2, 3])
def test_pk2d_basis_identities(k):
    n = sd.pk_size(k)
    nodes = [(i / k, j / k) for j in range(k + 1) for i in range(k + 1 - j)]       # i fastest
    for m, (x, y) in enumerate(nodes):
        v = sd.pk_basis(k, x, y)
        assert len(v) == n and np.allclose(v, np.eye(n)[m], atol=1e-13)            # Lagrange property in the stated numbering
    rng = np.random.default_rng(k)
    for _ in range(5):
        x, y = rng.uniform(0.05, 0.45, 2)
        v, g = sd.pk_basis(k, x, y), np.array(sd.pk_gradients(k, x, y))
        assert abs(sum(v) - 1.0) < 1e-13 and np.abs(g.sum(axis=0)).max() < 1e-12   # partition of unity
        h = 1e-6
        fd = np.array([[(a - b) / (2 * h) for a, b in zip(sd.pk_basis(k, x + h, y), sd.pk_basis(k, x - h, y))],
                       [(a - b) / (2 * h) for a, b in zip(sd.pk_basis(k, x, y + h), sd.pk_basis(k, x, y - h))]]).T
        assert np.abs(fd - g).max() < 1e-7
        # reproduces polynomials of degree k: sum_n p(node_n) phi_n = p
        p = lambda a, b: (1.0 + 2.0 * a - b) ** k
        assert abs(sum(p(*nodes[m]) * v[m] for m in range(n)) - p(x, y)) < 1e-12
    if k == 1:
        assert np.allclose(sd.pk_basis(1, 0.2, 0.3), [0.5, 0.2, 0.3])               # 1-x-y, x, y: the P1 basis of um_oracle


# a horizontal interface y = 1: Stokes cell above (inside), Darcy cell below (outside); edge from (0,1) to (2,1)
STOKES = [(0.0, 1.0), (2.0, 1.0), (0.5, 2.5)]      # face 0 = {v0, v1} is the interface, outer normal (0,-1)
DARCY = [(2.0, 1.0), (0.0, 1.0), (1.2, -0.7)]


def test_row_sums_of_the_three_terms_in_closed_form():
    P = sd.Params(gravity=9.81, alpha=0.7, viscosity=2e-3, porosity=0.4, gamma=1.5, density=1e3, kabs=((1e-7, 0.0), (0.0, 3e-7)))
    zero_v = [[0.0] * 6, [0.0] * 6]
    # (1) Darcy rows: -gamma porosity (u.n) psi_i; constant u = (0.3, -2): u.n = 2, partition of unity -> sum = -gamma porosity 2 |e|
    v = [sd.pk_interpolate(2, STOKES, lambda x, y: 0.3), sd.pk_interpolate(2, STOKES, lambda x, y: -2.0)]
    r_v, r_d = sd.alpha_coupling(P, STOKES, 0, DARCY, v, [0.0] * 10)
    assert abs(sum(r_d) - (-P.gamma * P.porosity * 2.0 * 2.0)) < 1e-12
    # (2) normal stress rows with u = 0, phi = 5 (constant head): sum_i r[d][i] = -rho g (5 - 1) n_d |e|; tangential term: K grad(phi) = 0
    head = sd.pk_interpolate(3, DARCY, lambda x, y: 5.0)
    r_v, r_d = sd.alpha_coupling(P, STOKES, 0, DARCY, zero_v, head)
    assert abs(sum(r_v[1]) - (-P.density * P.gravity * 4.0 * (-1.0) * 2.0)) < 1e-8 and abs(sum(r_v[0])) < 1e-9
    assert np.abs(r_d).max() == 0.0
    # (3) Beavers-Joseph rows with phi = 0 and the tangential velocity u = (1.5, 0): alpha sqrt(d)/sqrt(tr Pi) u_t |e| on component 0
    v = [sd.pk_interpolate(2, STOKES, lambda x, y: 1.5), [0.0] * 6]
    r_v, r_d = sd.alpha_coupling(P, STOKES, 0, DARCY, v, [0.0] * 10)
    tracePi = (1e-7 + 3e-7) * P.viscosity / P.gravity / P.density
    bj = P.alpha * math.sqrt(2.0) / math.sqrt(tracePi) * 1.5 * 2.0
    grav = -P.density * P.gravity * (0.0 - 1.0) * 2.0                              # -rho g (phi - z) n_d |e| with phi = 0, z = 1
    assert abs(sum(r_v[0]) - bj) < 1e-9 * bj and abs(sum(r_v[1]) - grav * (-1.0)) < 1e-8 and np.abs(r_d).max() < 1e-15


def test_hydrostatic_state_and_the_loop_bound_quirk():
    """u = 0, phi = z (head equal to the elevation): no flow, no normal stress jump.  The intended operator (gradients of all ten
    head functions transformed) returns exactly zero rows; the operator AS WRITTEN (loop over vsize = 6 of the dsize = 10 gradients,
    stokesdarcycouplingoperator.hh:194-195) sees a truncated grad(phi) whose tangential part drives the Beavers-Joseph rows."""
    P = sd.Params()
    zero_v = [[0.0] * 6, [0.0] * 6]
    head = sd.pk_interpolate(3, DARCY, lambda x, y: y)
    r_v, r_d = sd.alpha_coupling(P, STOKES, 0, DARCY, zero_v, head, quirk=False)
    assert np.abs(np.array(r_v)).max() < 1e-9 and np.abs(r_d).max() == 0.0
    # whether the truncation shows depends on which local edge of the Darcy cell the interface is: with the interface on local
    # face 0 the lost functions do not contribute a tangential gradient there, on the other faces they do
    r_vq, r_dq = sd.alpha_coupling(P, STOKES, 0, DARCY, zero_v, head, quirk=True)
    assert np.abs(np.array(r_vq)).max() < 1e-9
    for darcy in ([DARCY[2], DARCY[0], DARCY[1]], [DARCY[1], DARCY[2], DARCY[0]]):
        head = sd.pk_interpolate(3, darcy, lambda x, y: y)
        ok = sd.alpha_coupling(P, STOKES, 0, darcy, zero_v, head, quirk=False)
        assert np.abs(np.array(ok[0])).max() < 1e-9 and np.abs(ok[1]).max() == 0.0
        r_vq, r_dq = sd.alpha_coupling(P, STOKES, 0, darcy, zero_v, head, quirk=True)
        assert np.abs(np.array(r_vq[0])).max() > 0.5 and np.abs(r_dq).max() == 0.0  # tangential (x) rows are driven, O(1)
    # with a head whose last four coefficients vanish the two agree
    darcy = [DARCY[2], DARCY[0], DARCY[1]]
    head6 = list(sd.pk_interpolate(3, darcy, lambda x, y: y)[:6]) + [0.0] * 4
    a = sd.alpha_coupling(P, STOKES, 0, darcy, zero_v, head6, quirk=True)
    b = sd.alpha_coupling(P, STOKES, 0, darcy, zero_v, head6, quirk=False)
    assert np.allclose(a[0], b[0], rtol=0, atol=1e-12 * np.abs(b[0]).max())


def _dg_mesh():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import um_problems as ump
    return ump.square_mesh(3, 2, seed=5)


@pytest.mark.parametrize("k", [1, 2, 3])
def test_dg_on_triangles_is_consistent_symmetric_and_coercive(k):
    """ConvectionDiffusionDG on triangles with a piecewise constant tensor (the Darcy subproblem of stokesdarcy2): a polynomial of
    degree k with continuous normal flux solves the scheme exactly (zero residual with f = -div(A grad u), weak Dirichlet data g = u
    and Neumann data j = -A grad u . n on one side); the SIPG matrix is symmetric and positive definite; NIPG is not symmetric."""
    xy, conn = _dg_mesh()
    A0 = np.array([[2.0, 0.3], [0.3, 1.0]])
    A = [A0 for _ in range(conn.shape[0])]
    c = {1: (0, 0, 0, 0), 2: (0.7, -0.4, 0, 0), 3: (0.7, -0.4, 0.5, -0.3)}[k]
    u = lambda x, y: 1.0 + 2.0 * x - y + c[0] * x * x + c[1] * x * y + c[2] * x ** 3 + c[3] * x * y * y
    gu = lambda x, y: np.array([2.0 + 2 * c[0] * x + c[1] * y + 3 * c[2] * x * x + c[3] * y * y, -1.0 + c[1] * x + 2 * c[3] * x * y])
    hess = lambda x, y: np.array([[2 * c[0] + 6 * c[2] * x, c[1] + 2 * c[3] * y], [c[1] + 2 * c[3] * y, 2 * c[3] * x]])
    f = lambda x, y: -float(np.sum(A0 * hess(x, y)))
    bctype = lambda x, y: 'N' if x > 1.0 - 1e-9 else 'D'                           # Neumann on the right side, outer normal (1, 0)
    j = lambda x, y: -float((A0 @ gu(x, y))[0])
    dg = sd.DGTriangles(xy, conn, k, A, f, bctype, u, j, alpha=4.0)
    x = dg.interpolate(u)
    assert np.abs(dg.residual(x)).max() < 1e-10
    J, r0 = dg.jacobian()
    assert np.abs(J - J.T).max() < 1e-11 * np.abs(J).max()
    assert np.linalg.eigvalsh(0.5 * (J + J.T)).min() > 0
    assert np.abs(np.linalg.solve(J, -r0) - x).max() < 1e-9                       # the discrete solution is the polynomial
    nipg = sd.DGTriangles(xy, conn, k, A, f, bctype, u, j, alpha=4.0, theta=1.0)
    assert np.abs(nipg.residual(x)).max() < 1e-10
    Jn, _ = nipg.jacobian()
    assert np.abs(Jn - Jn.T).max() > 1e-3


def test_dg_harmonic_weights_with_a_jumping_tensor():
    """Two tensors across x = 0.5 (here the mesh line closest to it): a piecewise linear u with continuous normal flux
    A_l u_l' = A_r u_r' is reproduced (consistency with weightsOn), and the SIPG matrix stays symmetric."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import um_problems as ump
    xy, conn = ump.square_mesh(4, 2, seed=2, jitter=0.0)
    kl, kr = 1.0, 5.0
    left = xy[conn].mean(axis=1)[:, 0] < 0.5
    A = [np.eye(2) * (kl if l else kr) for l in left]
    u = lambda x, y: x / kl if x <= 0.5 else 0.5 / kl + (x - 0.5) / kr              # flux 1 everywhere
    bctype = lambda x, y: 'D' if (x < 1e-9 or x > 1 - 1e-9) else 'N'
    dg = sd.DGTriangles(xy, conn, 1, A, lambda x, y: 0.0, bctype, u, lambda x, y: 0.0, alpha=3.0)
    # interpolate per cell with the cell's own branch (u is only piecewise smooth)
    x = np.concatenate([[(p[0] / kl if l else 0.5 / kl + (p[0] - 0.5) / kr) for p in xy[conn[e]]] for e, l in enumerate(left)])
    assert np.abs(dg.residual(x)).max() < 1e-12
    J, _ = dg.jacobian()
    assert np.abs(J - J.T).max() < 1e-12 * np.abs(J).max()


def test_taylor_hood_stokes_manufactured_solution_and_symmetry():
    """u = (y^2, x^2) (divergence free, in P2), p = 1 + x - 2 y (in P1), f = -mu Laplace(u) + grad(p): every pressure row and every
    velocity row of a DOF off the boundary vanishes at the interpolant; the operator is symmetric; with the boundary velocity fixed and
    the pressure pinned at one vertex the discrete solution IS the interpolant."""
    xy, conn = _dg_mesh()
    mu = 0.7
    u = lambda x, y: (y * y, x * x)
    p = lambda x, y: 1.0 + x - 2.0 * y
    f = lambda x, y: (-mu * 2.0 + 1.0, -mu * 2.0 - 2.0)
    th = sd.StokesTaylorHood(xy, conn, mu=mu, f=f)
    x = th.interpolate(u, p)
    r = th.residual(x)
    bnd = th.boundary_velocity_dofs()
    free = np.ones(th.ndof, dtype=bool); free[bnd] = False
    assert np.abs(r[free]).max() < 1e-12 and np.abs(r[~free]).max() > 1e-2
    J, r0 = th.jacobian()
    assert np.abs(J - J.T).max() < 1e-12 * np.abs(J).max()
    # saddle point: velocity block positive definite on the free velocity DOFs, pressure block zero
    nv2 = 2 * th.n2
    vfree = free.copy(); vfree[nv2:] = False
    assert np.linalg.eigvalsh(J[np.ix_(vfree, vfree)]).min() > 0 and np.abs(J[nv2:, nv2:]).max() == 0.0
    # solve with Dirichlet velocity on the boundary and the pressure fixed at vertex 0
    fixed = ~free; fixed[nv2] = True
    A = J.copy(); b = -r0.copy()
    A[fixed] = 0.0; A[fixed, fixed] = 1.0; b[fixed] = x[fixed]
    assert np.abs(np.linalg.solve(A, b) - x).max() < 1e-9
