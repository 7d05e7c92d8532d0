"""Pins of oracle/um_sd_oracle.py, the CPU restatement of the reference's Stokes-Darcy path (SURVEY §8 a5 + a12:
test/stokesdarcy2.cc, test/stokesdarcycouplingoperator.hh, multi-child SubProblemLocalFunctionSpace).  The reference holds no
expected values ("parity unpinned"), so these are known answers: orthonormality of the OPB tables, numbering and pattern against
independent constructions, a manufactured Stokes solution, consistency of the DG scheme, closed-form coupling sums, FD = exact
derivative of the affine coupling residual."""
import os
import sys
from math import factorial

import numpy as np
import pytest
import scipy.sparse as sps

import sd_problems as sdp

capi = sdp.capi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_opb_tables  # noqa: E402
import um_sd_oracle as sdo  # noqa: E402


def test_opb_header_is_what_the_generator_writes():
    with open(os.path.join(ROOT, "dune-multidomain_b200", "csrc", "opb_tables.h")) as fp:
        assert fp.read() == gen_opb_tables.header_text()


@pytest.mark.parametrize("k", [1, 2, 3])
def test_opb_is_orthonormal_on_the_reference_triangle_and_lower_triangular(k):
    L, mon = gen_opb_tables.opb_matrix(k)
    n = len(mon)
    assert n == (k + 1) * (k + 2) // 2 and mon[:3] == [(0, 0), (1, 0), (0, 1)]
    assert all(L[i][j] == 0.0 for i in range(n) for j in range(i + 1, n)) and all(L[i][i] > 0 for i in range(n))
    G = np.zeros((n, n))
    for (x, y), w in sdo.tri_rule(2 * k):
        v, _ = sdo.opb_eval(k, x, y)
        G += w * np.outer(v, v)
    assert np.abs(G - np.eye(n)).max() < 1e-13
    assert abs(L[0][0] - np.sqrt(2.0)) < 1e-15                       # constant function with int_T phi^2 = 1 on |T| = 1/2
    h = 1e-6                                                         # gradients against central differences
    v, g = sdo.opb_eval(k, 0.3, 0.2)
    fx = (sdo.opb_eval(k, 0.3 + h, 0.2)[0] - sdo.opb_eval(k, 0.3 - h, 0.2)[0]) / (2 * h)
    fy = (sdo.opb_eval(k, 0.3, 0.2 + h)[0] - sdo.opb_eval(k, 0.3, 0.2 - h)[0]) / (2 * h)
    assert np.abs(g[:, 0] - fx).max() < 1e-6 and np.abs(g[:, 1] - fy).max() < 1e-6


def test_rules_are_exact():
    for order in range(1, 8):
        for a in range(order + 1):
            for b in range(order + 1 - a):
                exact = factorial(a) * factorial(b) / factorial(a + b + 2)
                num = sum(w * x ** a * y ** b for (x, y), w in sdo.tri_rule(order))
                assert abs(num - exact) < 1e-15
        assert abs(sum(w * t ** order for t, w in sdo.line_rule(order)) - 1.0 / (order + 1)) < 1e-15


@pytest.fixture(scope="module")
def small():
    xy, conn, groups = sdp.filtration_mesh(6, 4, seed=3)
    o = sdp.oracle()
    return o, sdp.setup(o, xy, conn, groups), xy, conn, groups


def test_numbering_against_an_independent_construction(small):
    o, H, xy, conn, groups = small
    stokes = groups == 0
    sv = np.unique(conn[stokes])
    edges = {}
    for e in range(conn.shape[0]):
        for a, b in ((0, 1), (0, 2), (1, 2)):
            edges.setdefault((min(conn[e, a], conn[e, b]), max(conn[e, a], conn[e, b])), len(edges))
    se = sorted({edges[(min(conn[e, a], conn[e, b]), max(conn[e, a], conn[e, b]))] for e in np.nonzero(stokes)[0] for a, b in ((0, 1), (0, 2), (1, 2))})
    n2 = len(sv) + len(se)
    nd = int((~stokes).sum()) * 10
    assert list(o.get_child_offsets()) == [0, n2, 2 * n2, 2 * n2 + len(sv), 2 * n2 + len(sv) + nd] and H["ndof"] == 2 * n2 + len(sv) + nd
    ptr, dofs = o.get_dofmap()
    assert np.array_equal(np.diff(ptr), np.where(stokes, 15, 10))
    e = int(np.nonzero(stokes)[0][0])
    d = dofs[ptr[e]:ptr[e + 1]]
    vrank = {v: i for i, v in enumerate(sv)}
    erank = {g: i for i, g in enumerate(se)}
    ed = [len(sv) + erank[edges[(min(conn[e, a], conn[e, b]), max(conn[e, a], conn[e, b]))]] for a, b in ((0, 1), (0, 2), (1, 2))]
    vv = [vrank[v] for v in conn[e]]
    p2 = [vv[0], ed[0], vv[1], ed[1], ed[2], vv[2]]
    assert list(d) == p2 + [n2 + i for i in p2] + [2 * n2 + i for i in vv]                  # SubProblemLocalFunctionSpace order v_0 | v_1 | p
    e = int(np.nonzero(~stokes)[0][5])
    rank = int((~stokes)[:e].sum())
    assert list(dofs[ptr[e]:ptr[e + 1]]) == [2 * n2 + len(sv) + rank * 10 + i for i in range(10)]
    assert H["ncon"] == 0                                                                   # StressNeumann everywhere: nothing constrained


def test_pattern_is_the_union_of_cliques_skeleton_and_coupling_links(small):
    o, H, xy, conn, groups = small
    rowptr, colidx = o.matrix_get_pattern(H["mat"])
    ptr, dofs = o.get_dofmap()
    rows, cols = [], []
    for e in range(conn.shape[0]):
        d = dofs[ptr[e]:ptr[e + 1]]
        rows += list(np.repeat(d, len(d))); cols += list(np.tile(d, len(d)))
        for f in range(3):
            nb = int(o.nbr[e, f])
            if nb < 0:
                continue
            if (groups[e] > 0) == (groups[nb] > 0) and groups[e] == 0:
                continue                                                                      # conforming spaces: no skeleton links
            dn = dofs[ptr[nb]:ptr[nb + 1]]                                                    # DG skeleton or coupling face: full blocks
            rows += list(np.repeat(d, len(dn))); cols += list(np.tile(dn, len(d)))
    A = sps.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(H["ndof"], H["ndof"]))
    A.sum_duplicates(); A.sort_indices()
    assert np.array_equal(A.indptr, rowptr) and np.array_equal(A.indices, colidx)


def _dense(o, H):
    rowptr, colidx = o.matrix_get_pattern(H["mat"])
    return sps.csr_matrix((o.matrix_get_values(H["mat"]), colidx, rowptr), shape=(H["ndof"], H["ndof"])).toarray()


def test_stokes_manufactured_solution_and_saddle_point_structure():
    """u = (y^2, x^2) is divergence free with grad u + grad u^T = [[0, 2(x+y)],[2(x+y), 0]], p = 3 - 2x + y:
    -mu Laplace(u) + grad p = (-2 mu - 2, -2 mu + 1) = f makes the Taylor-Hood interpolant the exact discrete solution."""
    xy, conn, groups = sdp.filtration_mesh(4, 4, seed=5, X=2.0, Y=2.0, y_if=0.0)            # everything is free flow
    th, da, cp = sdp.ini_params(stokes_bc=capi.BC_ALL_DIRICHLET)
    th.mu = 0.7
    th.f[0], th.f[1] = -2 * 0.7 - 2.0, -2 * 0.7 + 1.0
    o = sdp.oracle()
    H = sdp.setup(o, xy, conn, groups, params=(th, da, cp), darcy=False)
    assert H["ncon"] > 0
    x = np.zeros(H["ndof"])
    off = o.get_child_offsets()
    ptr, dofs = o.get_dofmap()
    for e in range(conn.shape[0]):
        p = xy[conn[e]]
        d = dofs[ptr[e]:ptr[e + 1]]
        nodes = [p[0], 0.5 * (p[0] + p[1]), p[1], 0.5 * (p[0] + p[2]), 0.5 * (p[1] + p[2]), p[2]]
        for i, q in enumerate(nodes):
            x[d[i]], x[d[6 + i]] = q[1] ** 2, q[0] ** 2
        for i in range(3):
            x[d[12 + i]] = 3.0 - 2.0 * p[i][0] + p[i][1]
    r = o.residual(H["go"], x, np.zeros(H["ndof"]))
    assert np.abs(r).max() < 1e-12
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    A = _dense(o, H)
    free = o.get_constraints() == 0
    Aff = A[np.ix_(free, free)]
    assert np.abs(Aff - Aff.T).max() < 1e-13                                                  # symmetric saddle-point operator
    pp = np.arange(off[2], off[3])
    assert np.abs(A[np.ix_(pp, pp)]).max() == 0.0                                             # zero pressure block (stored, structurally)
    vfree = free.copy(); vfree[off[2]:] = False
    assert np.linalg.eigvalsh(A[np.ix_(vfree, vfree)]).min() > 0                              # viscous block positive definite
    r2 = o.residual(H["go"], np.zeros(H["ndof"]), np.zeros(H["ndof"]))                        # affine: R(x) = A x + R(0) on free rows
    assert np.abs((A @ x + r2)[free]).max() < 1e-11


def test_darcy_dg_is_consistent_and_symmetric():
    """A linear head solves the SIPG scheme exactly (f = 0, constant tensor per layer would break it across the lens: one group)."""
    xy, conn, groups = sdp.filtration_mesh(4, 3, seed=2, X=3.0, Y=2.0, y_if=2.0)
    groups[:] = 1
    th, da, cp = sdp.ini_params(darcy_bc=capi.BCType(capi.BC_ALL_DIRICHLET))
    da.g = capi.Function(capi.FN_POLY2, 1.0, 2.0, -3.0, 0.0, 0.0)
    o = sdp.oracle()
    H = sdp.setup(o, xy, conn, groups, params=(th, da, cp), stokes=False)
    x = np.zeros(H["ndof"])
    ptr, dofs = o.get_dofmap()
    for e in range(conn.shape[0]):                                                           # L2 projection onto the orthonormal basis
        p = xy[conn[e]]
        c = np.zeros(10)
        for (lx, ly), w in sdo.tri_rule(4):
            q = p[0] + (p[1] - p[0]) * lx + (p[2] - p[0]) * ly
            c += w * (1.0 + 2.0 * q[0] - 3.0 * q[1]) * sdo.opb_eval(3, lx, ly)[0]
        x[dofs[ptr[e]:ptr[e + 1]]] = c
    r = o.residual(H["go"], x, np.zeros(H["ndof"]))
    assert np.abs(r).max() < 1e-12 * 1.16e-5 * 1e4
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    A = _dense(o, H)
    assert np.abs(A - A.T).max() < 1e-12 * np.abs(A).max() and np.linalg.eigvalsh(0.5 * (A + A.T)).min() > 0


def test_coupling_closed_forms_and_fd_equals_exact(small):
    o, H, xy, conn, groups = small
    th, da, cp = H["params"]
    cpd = o.cps[0]
    faces = [(e, f) for e in range(conn.shape[0]) for f in range(3) if o._fires(0, e, f)]
    assert len(faces) == 6                                                                    # the straight interface of the 6 x 4 mesh
    e, f = faces[0]
    nb = int(o.nbr[e, f])
    la, lb, pa, (dx, dy), fie, n = o._face(e, f)
    assert abs(n[0]) < 1e-14 and abs(n[1] + 1.0) < 1e-14                                      # free flow above: outer normal (0, -1)
    xs, xn = np.zeros(15), np.zeros(10)
    xs[0:6], xs[6:12] = 0.3, -2.0                                                             # constant velocity (Lagrange basis)
    rs, rn = o._sd_coupling(cpd, e, f, xs, xn)
    # Darcy rows: -gamma porosity (u.n) int psi_i; psi_0 = sqrt(2) is the constant function, u.n = 2: the total flux sits in row 0
    assert abs(rn[0] / np.sqrt(2.0) - (-cp.gamma * cp.porosity * 2.0 * fie)) < 1e-12 * fie
    # hydrostatic state of the intended operator (quirk off): u = 0, head = elevation -> no force on the free flow
    head = np.zeros(10)
    p = xy[conn[nb]]
    for (lx, ly), w in sdo.tri_rule(4):
        q = p[0] + (p[1] - p[0]) * lx + (p[2] - p[0]) * ly
        head += w * q[1] * sdo.opb_eval(3, lx, ly)[0]
    cp.quirk = 0
    rs, rn = o._sd_coupling(cpd, e, f, np.zeros(15), head)
    normal_rows = rs[6:12]
    assert np.abs(normal_rows).max() < 1e-9 * cp.density * cp.gravity and np.abs(rn).max() == 0.0
    cp.quirk = 1
    # finite differences (epsilon = 1e-2) of the affine residual = exact derivative, also with the quirk
    x = H["x0"]
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    A_fd = _dense(o, H)
    o.cps[0]["jac_mode"] = sdo.JAC_ANALYTIC
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    A_ex = _dense(o, H)
    o.cps[0]["jac_mode"] = sdo.JAC_FD
    assert np.abs(A_fd - A_ex).max() < 1e-9 * np.abs(A_ex).max()
    r0 = o.residual(H["go"], np.zeros(H["ndof"]), np.zeros(H["ndof"]))
    r1 = o.residual(H["go"], x, np.zeros(H["ndof"]))
    assert np.abs(A_ex @ x + r0 - r1).max() < 1e-10 * np.abs(r1).max()                         # the whole coupled residual is affine


@pytest.mark.parametrize("path", sdp.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_reproduces_golden(path):
    sdp.check_golden(sdp.oracle(), path)


def test_golden_fixtures_exist():
    assert len(sdp.golden_cases()) >= 3
