"""Pins of the unstructured-mesh oracle (oracle/um_oracle.py) — self-derived known answers, since the reference's tests hold no
expected values: independent numbering and adjacency constructions, the cotangent formula, closed-form mass and
coupling blocks, polynomial exactness of the rules, FD against the exact derivative, consistency of residual and Jacobian."""
import math

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import um_problems as ump

capi = ump.capi


def _csr(o, mat):
    rp, ci = o.matrix_get_pattern(mat)
    return sps.csr_matrix((o.matrix_get_values(mat), ci, rp), shape=(o.ndof, o.ndof))


@pytest.fixture(scope="module")
def case():
    xy, conn = ump.square_mesh(6, 5, seed=3)
    sd = ump.halves(xy, conn)
    return xy, conn, sd


def test_rules_are_exact():
    import um_oracle as uo
    for order in (1, 2, 3, 4):
        for a in range(order + 1):
            for b in range(order + 1 - a):
                exact = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
                got = sum(w * x ** a * y ** b for (x, y), w in uo.tri_rule(order))
                assert abs(got - exact) < 1e-14, (order, a, b)
    for order in range(0, 6):
        for a in range(order + 1):
            assert abs(sum(w * x ** a for x, w in uo.line_rule(order)) - 1.0 / (a + 1)) < 1e-15


def test_numbering_is_rank_of_the_subdomain_vertices(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd)
    off = o.get_child_offsets()
    ptr, dofs = o.get_dofmap()
    assert np.all(np.diff(ptr) == 3)                    # every cell lies in exactly one subdomain here
    for s in (0, 1):
        cells = np.nonzero((sd >> s) & 1)[0]
        verts = np.unique(conn[cells])                  # ascending host vertex index
        assert off[s + 1] - off[s] == verts.size
        rank = {v: i for i, v in enumerate(verts)}
        for e in cells:
            assert list(dofs[ptr[e]:ptr[e + 1]]) == [off[s] + rank[v] for v in conn[e]]
    assert H["ndof"] == off[-1]


def test_pattern_is_the_union_of_cell_cliques_and_coupling_links(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd)
    ptr, dofs = o.get_dofmap()
    n = H["ndof"]
    cell_dofs = dofs.reshape(-1, 3)
    B = sps.csr_matrix((np.ones(dofs.size), (np.repeat(np.arange(conn.shape[0]), 3), dofs)), shape=(conn.shape[0], n))
    A = (B.T @ B).tolil()
    # interface edges: cells of different subdomains sharing two vertices; the coupling fires from subdomain 0
    edge = {}
    for e in range(conn.shape[0]):
        for a, b in ((0, 1), (0, 2), (1, 2)):
            edge.setdefault(frozenset((conn[e, a], conn[e, b])), []).append(e)
    nfaces = 0
    for cells in edge.values():
        if len(cells) == 2 and sd[cells[0]] != sd[cells[1]]:
            nfaces += 1
            for i in cell_dofs[cells[0]]:
                for j in cell_dofs[cells[1]]:
                    A[i, j] = 1; A[j, i] = 1
    assert nfaces == 5
    A = A.tocsr(); A.sort_indices()
    rp, ci = o.matrix_get_pattern(H["mat"])
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)


def test_stiffness_is_the_cotangent_formula_and_mass_closed_form(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd, coupling=False, with_mass=True, bc=capi.BCType(capi.BC_ALL_NEUMANN))
    assert H["ncon"] == 0
    ptr, dofs = o.get_dofmap()
    n = H["ndof"]
    K, M = sps.lil_matrix((n, n)), sps.lil_matrix((n, n))
    for e in range(conn.shape[0]):
        p = xy[conn[e]]
        d = dofs[ptr[e]:ptr[e + 1]]
        area = 0.5 * abs((p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[1, 1] - p[0, 1]) * (p[2, 0] - p[0, 0]))
        for i in range(3):
            for j in range(3):
                M[d[i], d[j]] += area / 12.0 * (2.0 if i == j else 1.0)
            j, k = (i + 1) % 3, (i + 2) % 3                   # edge (j, k) is opposite to vertex i
            u, v = p[j] - p[i], p[k] - p[i]
            cot = np.dot(u, v) / abs(u[0] * v[1] - u[1] * v[0])
            K[d[j], d[k]] -= 0.5 * cot; K[d[k], d[j]] -= 0.5 * cot
            K[d[j], d[j]] += 0.5 * cot; K[d[k], d[k]] += 0.5 * cot
    x = np.zeros(n)
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    assert abs(_csr(o, H["mat"]) - K.tocsr()).max() < 1e-12 * abs(K).max()
    o.jacobian(H["go1"], H["mat"], x, weight=0.5, overwrite=True)
    assert abs(_csr(o, H["mat"]) - 0.5 * M.tocsr()).max() < 1e-14


def test_patch_test_and_coupling_vanishes_on_continuous_functions(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd, f=capi.Function(capi.FN_ZERO), bc=capi.BCType(capi.BC_ALL_DIRICHLET))
    lin = capi.Function(capi.FN_POLY2, 1.0, 2.0, -3.0, 0.0, 0.0)
    x = o.interpolate(list(H["sp"]), [lin, lin], np.zeros(H["ndof"]))
    r = o.residual(H["go"], x, np.zeros(H["ndof"]))
    off = o.get_child_offsets()
    for s in (0, 1):                                        # rows of vertices off the interface x = 0.5 and off the boundary
        verts = np.unique(conn[np.nonzero((sd >> s) & 1)[0]])
        inner = np.abs(xy[verts, 0] - 0.5) > 1e-9
        assert np.abs(r[off[s]:off[s + 1]][inner]).max() < 1e-13
    assert np.abs(r).max() > 1e-2                           # the one-sided flux on the interface rows is there


def test_cvcf_coupling_is_consistent_and_symmetric(case):
    """With ContinuousValueContinuousFlowCoupling (theta = 1: symmetric interior penalty) the two-subdomain problem is consistent:
    a globally linear function has zero residual on EVERY unconstrained row, interface rows included (the coupling's flux term
    cancels the one-sided boundary flux of the volume terms, its jump terms vanish) — which is not so without the coupling.  The
    exact-derivative Jacobian of the coupled problem is symmetric on the unconstrained block."""
    xy, conn, sd = case
    lin = capi.Function(capi.FN_POLY2, 1.0, 2.0, -3.0, 0.0, 0.0)
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd, cvcf=True, k=2.0, f=capi.Function(capi.FN_ZERO), bc=capi.BCType(capi.BC_ALL_DIRICHLET),
                  jac_mode=capi.JAC_ANALYTIC)
    x = o.interpolate(list(H["sp"]), [lin, lin], np.zeros(H["ndof"]))
    r = o.residual(H["go"], x, np.zeros(H["ndof"]))
    assert np.abs(r).max() < 1e-13
    o.jacobian(H["go"], H["mat"], x, overwrite=True)
    free = o.get_constraints() == 0
    A = _csr(o, H["mat"]).toarray()[np.ix_(free, free)]
    assert np.abs(A - A.T).max() < 1e-12 * np.abs(A).max()
    assert np.linalg.eigvalsh(0.5 * (A + A.T)).min() > 0           # penalty 2 / h_F is enough on this mesh
    o2 = ump.oracle()
    H2 = ump.setup(o2, xy, conn, sd, cvcf=True, k=2.0, f=capi.Function(capi.FN_ZERO), bc=capi.BCType(capi.BC_ALL_DIRICHLET))
    o2.jacobian(H2["go"], H2["mat"], x, overwrite=True)
    fd = _csr(o2, H2["mat"]).toarray()
    an = _csr(o, H["mat"]).toarray()
    assert 0 < np.abs(fd - an).max() < 1e-5 * np.abs(an).max()     # FD noise, and nothing more


def test_fd_coupling_jacobian_against_closed_form(case):
    xy, conn, sd = case
    k = 2.5
    mats = {}
    for mode in (capi.JAC_FD_BITMATCH, capi.JAC_ANALYTIC):
        o = ump.oracle()
        H = ump.setup(o, xy, conn, sd, jac_mode=mode, k=k, bc=capi.BCType(capi.BC_ALL_NEUMANN))
        o.jacobian(H["go"], H["mat"], H["x0"], overwrite=True)
        with_c = _csr(o, H["mat"])
        o2 = ump.oracle()
        H2 = ump.setup(o2, xy, conn, sd, coupling=False, bc=capi.BCType(capi.BC_ALL_NEUMANN))
        o2.jacobian(H2["go"], H2["mat"], H2["x0"], overwrite=True)
        mats[mode] = (with_c - _csr(o2, H2["mat"])).toarray()
        off = o.get_child_offsets()
        v2d = [o.children[0]["v2d"], o.children[1]["v2d"]]
    C = np.zeros_like(mats[capi.JAC_ANALYTIC])
    for v in np.nonzero(np.abs(xy[:, 0] - 0.5) < 1e-12)[0]:
        for w in np.nonzero(np.abs(xy[:, 0] - 0.5) < 1e-12)[0]:
            L = abs(xy[v, 1] - xy[w, 1])
            shared = [e for e in range(conn.shape[0]) if v in conn[e] and w in conn[e]]
            if v != w and len(shared) != 2:
                continue
            if v == w:                                      # both interface edges at v contribute |e| / 3
                m = sum(abs(xy[v, 1] - xy[u, 1]) / 3.0 for u in np.nonzero(np.abs(xy[:, 0] - 0.5) < 1e-12)[0]
                        if u != v and sum(1 for e in range(conn.shape[0]) if v in conn[e] and u in conn[e]) == 2)
            else:
                m = L / 6.0
            i0, j0, i1, j1 = off[0] + v2d[0][v], off[0] + v2d[0][w], off[1] + v2d[1][v], off[1] + v2d[1][w]
            C[i0, j0] += k * m; C[i0, j1] -= k * m; C[i1, j0] -= k * m; C[i1, j1] += k * m
    assert np.abs(mats[capi.JAC_ANALYTIC] - C).max() < 1e-14
    assert 0 < np.abs(mats[capi.JAC_FD_BITMATCH] - C).max() < 1e-6


def test_newton_step_solves_the_affine_problem(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd, jac_mode=capi.JAC_ANALYTIC)
    assert H["ncon"] > 0
    o.jacobian(H["go"], H["mat"], H["x0"], overwrite=True)
    r = o.residual(H["go"], H["x0"], np.zeros(H["ndof"]))
    z = spla.spsolve(_csr(o, H["mat"]).tocsc(), r)
    x = H["x0"] - z
    r1 = o.residual(H["go"], x, np.zeros(H["ndof"]))
    assert np.abs(r1).max() < 1e-11 * max(1.0, np.abs(r).max())
    con = o.get_constraints() == 1
    assert np.abs(x[con] - H["x0"][con]).max() < 1e-15      # Dirichlet values stay (unit rows, zero residual)


def test_shared_space_variant_runs(case):
    xy, conn, sd = case
    o = ump.oracle()
    H = ump.setup(o, xy, conn, sd, shared_space=True)
    assert H["ndof"] == xy.shape[0]
    o.jacobian(H["go"], H["mat"], H["x0"], overwrite=True)
    assert np.isfinite(o.matrix_get_values(H["mat"])).all()


@pytest.mark.parametrize("path", ump.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_oracle_reproduces_golden(path):
    ump.check_golden(ump.oracle(), path)


def test_golden_fixtures_exist():
    assert len(ump.golden_cases()) >= 3
