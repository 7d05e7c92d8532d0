"""Pins for the oracle's enriched (mortar) coupling path — SURVEY §8 a11, BASELINE configs[3] (test/testmortarpoisson.cc).
No GPU.  The reference ships no expected values, so the restatement is pinned by an independent NumPy construction of the
mortar numbering and of the sparsity pattern, closed-form blocks of the Jacobian, algebraic identities and SciPy solves."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

import orc
import problems

capi = orc.capi


def _setup(n, **kw):
    o = orc.Oracle()
    return o, problems.testmortarpoisson(o, n, **kw)


def test_sizes_of_the_shipped_testmortarpoisson():
    """refinement 5 (testmortarpoisson.cc:289): 32x32 cells, two 17x33 Q1 blocks, 33 mortar DOFs on the interface x0 = 0.5;
    `std::cout << multigfs.ordering().size()` (:386) prints 1155."""
    o, P = _setup((32, 32))
    assert list(o.get_child_offsets()) == [0, 17 * 33, 2 * 17 * 33, 2 * 17 * 33 + 33]
    assert P.ndof == 1155


@pytest.mark.parametrize("n", [(6, 4), (4, 3, 5)])
def test_mortar_numbering_and_cell_spaces(n):
    """couplinggfsordering.hh:150-200: mortar DOFs = vertices of the accepted intersections by ascending MultiDomainGrid vertex
    index, appended as the last child block; a cell's local space holds no mortar DOF (they are bound per intersection)."""
    o, P = _setup(n)
    dim = len(n)
    off = o.get_child_offsets()
    nv = np.array(n) + 1
    ivert = [v for v in range(int(np.prod(nv))) if v % nv[0] == n[0] // 2]          # vertices on the plane x0 = 0.5, ascending index
    assert off[3] - off[2] == len(ivert)
    ptr, dofs = o.get_dofmap()
    assert np.all(np.diff(ptr) == 2 ** dim) and dofs.max() < off[2]
    # pattern rows of the mortar block: row of vertex v holds the mortar DOFs of every interface face around v
    rp, ci = o.matrix_get_pattern(P.mat)
    coords = np.array([np.unravel_index(v, nv[::-1])[::-1] for v in ivert])
    for a, va in enumerate(ivert):
        cols = ci[rp[off[2] + a]:rp[off[2] + a + 1]]
        cc = cols[cols >= off[2]] - off[2]
        expect = [b for b in range(len(ivert)) if np.abs(coords[b] - coords[a]).max() <= 1]
        assert list(cc) == expect


@pytest.mark.parametrize("n,fe", [((6, 4), (capi.FE_Q1, capi.FE_Q1)), ((4, 4), (capi.FE_Q1, capi.FE_Q2)), ((4, 3, 2), (capi.FE_Q1, capi.FE_Q1))])
def test_pattern_is_union_of_volume_cliques_and_enriched_links(n, fe):
    """FullVolumePattern cliques + FullEnrichedCoupling{First,Second}SubProblemPattern (sc, cs, nc, cn) + FullEnrichedCouplingPattern
    (cc) over the interface faces (couplingutilities.hh:219-286) — and NO sn / ns links (the enriched coupling has no pattern_coupling)."""
    o, P = _setup(n, fe=fe)
    dim = len(n)
    ptr, dofs = o.get_dofmap()
    off = o.get_child_offsets()
    rows, cols = [], []
    for c in range(len(ptr) - 1):
        d = dofs[ptr[c]:ptr[c + 1]]
        rows.append(np.repeat(d, len(d))); cols.append(np.tile(d, len(d)))
    # interface faces: left cell (i0 = n0/2 - 1) -> right cell (i0 = n0/2); mortar vertex index among the plane's vertices
    nv = np.array(n) + 1
    plane = {v: k for k, v in enumerate(v for v in range(int(np.prod(nv))) if v % nv[0] == n[0] // 2)}
    other = [range(n[d]) for d in range(1, dim)]
    for idx in np.ndindex(*[len(r) for r in other]):
        cl = n[0] // 2 - 1 + n[0] * (idx[0] + (n[1] * idx[1] if dim == 3 else 0))
        cr = cl + 1
        verts = []
        for corner in range(2 ** (dim - 1)):
            v = [n[0] // 2] + [idx[t] + ((corner >> t) & 1) for t in range(dim - 1)]
            verts.append(plane[v[0] + nv[0] * (v[1] + (nv[1] * v[2] if dim == 3 else 0))])
        m = off[2] + np.array(verts)
        for d in (dofs[ptr[cl]:ptr[cl + 1]], dofs[ptr[cr]:ptr[cr + 1]]):
            rows += [np.repeat(d, len(m)), np.repeat(m, len(d))]; cols += [np.tile(m, len(d)), np.tile(d, len(m))]
        rows.append(np.repeat(m, len(m))); cols.append(np.tile(m, len(m)))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    ref = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(P.ndof, P.ndof))
    ref.sum_duplicates(); ref.sort_indices()
    rp, ci = o.matrix_get_pattern(P.mat)
    assert np.array_equal(rp, ref.indptr) and np.array_equal(ci, ref.indices)


def test_cc_block_closed_form_2d():
    """d r_c,i / d lambda_j = 2 gamma int mu_i mu_j from each side (testmortarpoisson.cc:236-239), gamma = gamma_0 / h: per face
    2 * 2 gamma_0 * [[1/3, 1/6], [1/6, 1/3]] (the 2-point Gauss rule of qorder 2 integrates mu_i mu_j exactly)."""
    n, g0 = (8, 6), 1.75
    o, P = _setup(n, gamma_0=g0, jac_mode=capi.JAC_ANALYTIC)
    o.jacobian(P.go, P.mat, np.zeros(P.ndof), overwrite=True)
    A = o.csr(P.mat)
    off = o.get_child_offsets()
    cc = A[off[2]:, off[2]:].toarray()
    nm = n[1] + 1
    ref = np.zeros((nm, nm))
    for e in range(n[1]):
        ref[e:e + 2, e:e + 2] += 4 * g0 * np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    assert np.abs(cc - ref).max() < 1e-14


@pytest.mark.parametrize("n,fe", [((8, 8), (capi.FE_Q1, capi.FE_Q1)), ((4, 6), (capi.FE_Q2, capi.FE_Q1)), ((4, 4, 3), (capi.FE_Q1, capi.FE_Q1))])
def test_enriched_jacobian_identities(n, fe):
    """(i) the FD Jacobian (couplingutilities.hh:292-484, eps = 1e-8) is within FD noise of the exact one; (ii) the exact one is
    symmetric on the free rows (ss, sc = cs^T, cc are symmetric forms); (iii) R is affine: R(x) = J x + R(0) on the free rows."""
    o, P = _setup(n, fe=fe)
    o2, P2 = _setup(n, fe=fe, jac_mode=capi.JAC_ANALYTIC)
    x = np.random.default_rng(0).uniform(-1, 1, P.ndof)
    o.jacobian(P.go, P.mat, x, overwrite=True); o2.jacobian(P2.go, P2.mat, x, overwrite=True)
    A, B = o.csr(P.mat), o2.csr(P2.mat)
    assert abs(A - B).max() < 2e-6 * abs(B).max()
    free = ~o.get_constraints().astype(bool)
    Bf = B[free][:, free]
    assert abs(Bf - Bf.T).max() < 1e-13 * abs(B).max()
    r0, rx = np.zeros(P.ndof), np.zeros(P.ndof)
    o.residual(P.go, np.zeros(P.ndof), r0); o.residual(P.go, x, rx)
    d = (B @ x + r0 - rx)[free]
    assert np.abs(d).max() < 1e-12 * np.abs(rx).max()


def test_mortar_stationary_solve_against_scipy():
    """StationaryLinearProblemSolver with ISTLBackend_SEQ_BCGS_SSOR(5000) and reduction 1e-10 (testmortarpoisson.cc:456-463): the
    saddle-point-like system (indefinite: the mortar block couples with both signs) is solved by BiCGSTAB + SSOR(3); the
    multiplier reproduces the trace of the left solution up to the discretisation error."""
    n = (16, 16)
    o, P = _setup(n)
    u = np.zeros(P.ndof); o.interpolate(P.sps, P.gfn, u)
    r = np.zeros(P.ndof); o.residual(P.go, u, r)
    o.jacobian(P.go, P.mat, u, overwrite=True)
    z = np.zeros(P.ndof)
    st = o.solve(P.mat, r, z, reduction=1e-10)
    assert st.converged
    zs = sla.spsolve(o.csr(P.mat).tocsc(), r)
    assert np.abs(z - zs).max() < 1e-7 * np.abs(zs).max()
    un = u - zs
    off = o.get_child_offsets()
    ul = un[off[0]:off[1]].reshape(n[1] + 1, n[0] // 2 + 1)[:, -1]
    assert np.abs(ul - un[off[2]:]).max() < 5e-3


# ---------------------------------------------------------------------------------------------------------------------
# PowerCouplingGridFunctionSpace<CouplingGFS, 2, VBE> (test/testpowermortarpoisson.cc:384-386): two copies of the mortar
# space, block after block (LexicographicOrderingTag), the operator loops over the blocks (:186-189)
# ---------------------------------------------------------------------------------------------------------------------

def test_sizes_of_the_shipped_testpowermortarpoisson():
    """refinement 5 (testpowermortarpoisson.cc:296): two 17x33 Q1 blocks and two mortar blocks of 33 DOFs each."""
    o, P = _setup((32, 32), ncomp=2)
    assert list(o.get_child_offsets()) == [0, 561, 1122, 1155, 1188]
    assert P.ndof == 1188


@pytest.mark.parametrize("n,fe", [((6, 4), (capi.FE_Q1, capi.FE_Q1)), ((4, 4), (capi.FE_Q2, capi.FE_Q1)), ((4, 3, 2), (capi.FE_Q1, capi.FE_Q1))])
def test_power_pattern_from_the_single_mortar_pattern(n, fe):
    """The local coupling space of an intersection is the concatenation of the component blocks' spaces, and every enriched pattern
    is full over it (couplingutilities.hh:219-286): with S the single-mortar pattern in blocks [ee ec; ce cc], the power pattern is
    [ee ec ec; ce cc cc; ce cc cc] — including the cross links between the two multiplier blocks, which the operator never fills."""
    o1, P1 = _setup(n, fe=fe)
    o2, P2 = _setup(n, fe=fe, ncomp=2)
    off = o1.get_child_offsets()
    ne, nm = off[2], off[3] - off[2]
    assert list(o2.get_child_offsets()) == [off[0], off[1], off[2], ne + nm, ne + 2 * nm]
    rp, ci = o1.matrix_get_pattern(P1.mat)
    S = sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(ne + nm, ne + nm))
    ee, ec, ce, cc = S[:ne, :ne], S[:ne, ne:], S[ne:, :ne], S[ne:, ne:]
    ref = sp.bmat([[ee, ec, ec], [ce, cc, cc], [ce, cc, cc]], format="csr")
    ref.sort_indices()
    rp2, ci2 = o2.matrix_get_pattern(P2.mat)
    assert np.array_equal(rp2, ref.indptr) and np.array_equal(ci2, ref.indices)


@pytest.mark.parametrize("n,fe", [((8, 6), (capi.FE_Q1, capi.FE_Q1)), ((4, 4), (capi.FE_Q1, capi.FE_Q2)), ((4, 3, 3), (capi.FE_Q1, capi.FE_Q1))])
def test_power_jacobian_from_the_single_mortar_jacobian(n, fe):
    """Each component block runs the whole body of alpha_generic, so (exact derivative, free rows): the element-element coupling
    terms are added once per block (ee_power = vol + 2 (ee_single - vol)), every block has the single space's ec / ce / cc blocks,
    and the stored cross links cc_01, cc_10 stay zero."""
    o1, P1 = _setup(n, fe=fe, jac_mode=capi.JAC_ANALYTIC)
    o2, P2 = _setup(n, fe=fe, jac_mode=capi.JAC_ANALYTIC, ncomp=2)
    gov = o1.gridoperator_create(P1.sps, [])             # the two Poisson subproblems without the coupling
    matv = o1.matrix_create([gov])
    z1, z2 = np.zeros(P1.ndof), np.zeros(P2.ndof)
    o1.jacobian(P1.go, P1.mat, z1, overwrite=True); o1.jacobian(gov, matv, z1, overwrite=True); o2.jacobian(P2.go, P2.mat, z2, overwrite=True)
    A1, V, A2 = o1.csr(P1.mat), o1.csr(matv), o2.csr(P2.mat)
    off = o1.get_child_offsets()
    ne, nm = off[2], off[3] - off[2]
    free = ~o1.get_constraints().astype(bool)[:ne]
    scale = abs(A1).max()
    ee = (V[:ne, :ne] + 2 * (A1[:ne, :ne] - V[:ne, :ne]))
    assert abs((A2[:ne, :ne] - ee)[free]).max() < 1e-13 * scale
    for b in range(2):
        lo, hi = ne + b * nm, ne + (b + 1) * nm
        assert abs((A2[:ne, lo:hi] - A1[:ne, ne:])[free]).max() < 1e-13 * scale
        assert abs(A2[lo:hi, :ne] - A1[ne:, :ne]).max() < 1e-13 * scale
        assert abs(A2[lo:hi, lo:hi] - A1[ne:, ne:]).max() < 1e-13 * scale
    assert abs(A2[ne:ne + nm, ne + nm:]).max() == 0.0 and abs(A2[ne + nm:, ne:ne + nm]).max() == 0.0
    assert A2[ne:ne + nm, ne + nm:].nnz > 0              # ... but they are stored (full pattern over the power space)


@pytest.mark.parametrize("n", [(8, 8), (4, 4, 3)])
def test_power_fd_jacobian_and_affine_residual(n):
    """FD Jacobian of the power operator within FD noise of the exact one; R(x) = J x + R(0) on the free rows; and with equal
    multiplier blocks the residual's element rows equal vol + 2 (single - vol)."""
    o, P = _setup(n, ncomp=2)
    oa, Pa = _setup(n, ncomp=2, jac_mode=capi.JAC_ANALYTIC)
    x = np.random.default_rng(1).uniform(-1, 1, P.ndof)
    o.jacobian(P.go, P.mat, x, overwrite=True); oa.jacobian(Pa.go, Pa.mat, x, overwrite=True)
    A, B = o.csr(P.mat), oa.csr(Pa.mat)
    assert abs(A - B).max() < 2e-6 * abs(B).max()
    free = ~o.get_constraints().astype(bool)
    r0, rx = np.zeros(P.ndof), np.zeros(P.ndof)
    o.residual(P.go, np.zeros(P.ndof), r0); o.residual(P.go, x, rx)
    assert np.abs((B @ x + r0 - rx)[free]).max() < 1e-12 * np.abs(rx).max()
    o1, P1 = _setup(n)
    gov = o1.gridoperator_create(P1.sps, [])
    off = o1.get_child_offsets()
    ne, nm = off[2], off[3] - off[2]
    x1 = x[:ne + nm]
    x2 = np.concatenate([x1, x1[ne:]])
    r1, rv, r2 = np.zeros(P1.ndof), np.zeros(P1.ndof), np.zeros(P.ndof)
    o1.residual(P1.go, x1, r1); o1.residual(gov, x1, rv); o.residual(P.go, x2, r2)
    f1 = free[:ne]
    assert np.abs((r2[:ne] - (rv[:ne] + 2 * (r1[:ne] - rv[:ne])))[f1]).max() < 1e-13 * np.abs(r1).max()
    assert np.abs(r2[ne:ne + nm] - r1[ne:]).max() < 1e-13 * np.abs(r1).max() and np.array_equal(r2[ne:ne + nm], r2[ne + nm:])
