"""Unstructured-mesh test problems, driven identically through the product binding (capi.Mdb) and the Python oracle
(oracle/um_oracle.UMOracle): two-subdomain Poisson on a triangle mesh with P1 spaces and ProportionalFlowCoupling — the
operators of test/testpoisson.cc / test/proportionalflowcoupling.hh on the grid class of test/stokesdarcy2.cc:575-604."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = importlib.import_module("dune-multidomain_b200").capi


sys.path.insert(0, os.path.join(ROOT, "oracle"))
import um_oracle  # noqa: E402  (test infrastructure)


def oracle():
    return um_oracle.UMOracle()


def square_mesh(nx, ny, seed=0, jitter=0.15, shuffle=True, lower=(0.0, 0.0), upper=(1.0, 1.0)):
    """nx x ny squares cut into two triangles each (alternating diagonals), interior vertices moved by `jitter` mesh widths,
    boundary vertices only along the boundary; vertex order, cell order and the local vertex order of every cell are shuffled
    (seeded) so that no numbering coincides with a lattice."""
    rng = np.random.default_rng(seed)
    hx, hy = (upper[0] - lower[0]) / nx, (upper[1] - lower[1]) / ny
    X, Y = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    xy = np.stack([lower[0] + X.ravel() * hx, lower[1] + Y.ravel() * hy], axis=1)
    if jitter > 0:
        inner_x = (X.ravel() > 0) & (X.ravel() < nx)
        inner_y = (Y.ravel() > 0) & (Y.ravel() < ny)
        # keep the vertical mid line straight when nx is even: it is the subdomain interface of the tests
        straight = (nx % 2 == 0) & (X.ravel() == nx // 2)
        xy[:, 0] += np.where(inner_x & ~straight, rng.uniform(-jitter, jitter, xy.shape[0]) * hx, 0.0)
        xy[:, 1] += np.where(inner_y, rng.uniform(-jitter, jitter, xy.shape[0]) * hy, 0.0)
    conn = []
    vid = lambda i, j: j * (nx + 1) + i
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            if (i + j) % 2 == 0:
                conn += [[a, b, d], [a, d, c]]
            else:
                conn += [[a, b, c], [b, d, c]]
    conn = np.array(conn, dtype=np.int64)
    if shuffle:
        pv = rng.permutation(xy.shape[0])                # new index of old vertex v = pv[v]
        xy2 = np.empty_like(xy); xy2[pv] = xy
        conn = pv[conn]
        conn = conn[rng.permutation(conn.shape[0])]
        for e in range(conn.shape[0]):
            conn[e] = np.roll(conn[e], rng.integers(0, 3))
            if rng.integers(0, 2):
                conn[e] = conn[e][::-1]                  # orientation is free (|det J|)
        xy = xy2
    return np.ascontiguousarray(xy), np.ascontiguousarray(conn)


def halves(xy, conn, axis=0, threshold=0.5):
    """subdomain 1 where the cell centroid has x[axis] > threshold, else 0 (test/testpoisson.cc:152-158)."""
    c = xy[conn].mean(axis=1)
    return np.where(c[:, axis] > threshold, 2, 1).astype(np.uint8)


def setup(lib, xy, conn, sd, *, shared_space=False, jac_mode=capi.JAC_FD_BITMATCH, k=2.5, orders=(4, 2), f=None, with_mass=False,
          bc=None, coupling=True, cvcf=False):
    """Spaces, subproblems, coupling, constraints, linearisation point.  Returns a dict of handles."""
    lib.mesh_unstructured(xy, conn)
    lib.subdomains(sd)
    if shared_space:
        c0 = c1 = lib.add_space(capi.FE_P1, -1)
    else:
        c0, c1 = lib.add_space(capi.FE_P1, 0), lib.add_space(capi.FE_P1, 1)
    bc = capi.BCType(capi.BC_TESTPOISSON) if bc is None else bc
    f = capi.Function(capi.FN_POLY2, 1.0, 2.0, -3.0, 0.0, 4.0) if f is None else f
    P = []
    for o in orders:
        p = capi.PoissonParams()
        p.f, p.bc, p.j, p.quad_order = f, bc, capi.Function(capi.FN_TESTPOISSON_J), o
        P.append(p)
    sp0 = lib.add_subproblem(capi.OP_POISSON, P[0], capi.COND_EQUALITY, 1, [c0])
    sp1 = lib.add_subproblem(capi.OP_POISSON, P[1], capi.COND_EQUALITY, 2, [c1])
    H = dict(sp=(sp0, sp1), P=P)
    if with_mass:
        l2 = capi.L2Params(); l2.quad_order = 2
        H["mass_sp"] = (lib.add_subproblem(capi.OP_L2, l2, capi.COND_EQUALITY, 1, [c0]),
                        lib.add_subproblem(capi.OP_L2, l2, capi.COND_EQUALITY, 2, [c1]))
    cps = []
    if coupling and cvcf:                                    # the coupling of test/testpoisson.cc:234 (intorder 4)
        cv = capi.CVCFParams(); cv.quad_order, cv.intensity = 4, k
        cps = [lib.add_coupling(capi.OP_CVCF_COUPLING, cv, sp0, sp1, jac_mode)]
    elif coupling:
        pf = capi.PropFlowParams(); pf.intensity = k
        cps = [lib.add_coupling(capi.OP_PROPFLOW_COUPLING, pf, sp0, sp1, jac_mode)]
    H["ndof"] = lib.finalize()
    H["ncon"] = lib.constraints_assemble([sp0, sp1], [bc, bc])
    H["go"] = lib.gridoperator_create([sp0, sp1], cps)
    gos = [H["go"]]
    if with_mass:
        H["go1"] = lib.gridoperator_create(list(H["mass_sp"]), [])
        gos.append(H["go1"])
    H["mat"] = lib.matrix_create(gos)
    x = np.zeros(H["ndof"])
    lib.interpolate([sp0, sp1], [capi.gauss(), capi.gauss()], x)
    H["x0"] = x
    return H


# ---- Gmsh 2.x ASCII: a writer for fixtures and an independent Python reading of the reader's rules ---------------------------
def write_gmsh(path, xy, conn, groups, with_boundary=True, node_id_offset=1, extra_nodes=0):
    """Triangles (type 2) with their physical group as first tag, optionally the boundary edges as lines (type 1, group 7) in
    front of them and a point element (type 15), node ids starting at `node_id_offset`, `extra_nodes` unused nodes in front."""
    edges = {}
    for e in range(conn.shape[0]):
        for a, b in ((0, 1), (0, 2), (1, 2)):
            edges.setdefault((min(conn[e, a], conn[e, b]), max(conn[e, a], conn[e, b])), []).append(e)
    lines = [k for k, v in sorted(edges.items()) if len(v) == 1] if with_boundary else []
    nid = lambda v: int(v) + node_id_offset + extra_nodes
    with open(path, "w") as fp:
        fp.write("$MeshFormat\n2.1 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % (xy.shape[0] + extra_nodes))
        for i in range(extra_nodes):
            fp.write("%d %r %r 0\n" % (node_id_offset + i, -1.0 - i, -1.0))
        for v in range(xy.shape[0]):
            fp.write("%d %r %r 0\n" % (nid(v), float(xy[v, 0]), float(xy[v, 1])))
        fp.write("$EndNodes\n$Elements\n%d\n" % (1 + len(lines) + conn.shape[0]))
        k = 1
        fp.write("%d 15 3 0 1 0 %d\n" % (k, nid(0))); k += 1
        for a, b in lines:
            fp.write("%d 1 3 7 1 0 %d %d\n" % (k, nid(a), nid(b))); k += 1
        for e in range(conn.shape[0]):
            fp.write("%d 2 3 %d %d 0 %d %d %d\n" % (k, groups[e], 10 + groups[e], nid(conn[e, 0]), nid(conn[e, 1]), nid(conn[e, 2]))); k += 1
        fp.write("$EndElements\n")


def read_gmsh(path):
    """[UPSTREAM dune-grid GmshReader]: nodes used by lines / triangles become vertices in order of first appearance."""
    tok = open(path).read().split()
    i = tok.index("$Nodes") + 1
    n = int(tok[i]); i += 1
    nodes = {}
    for _ in range(n):
        nodes[int(tok[i])] = (float(tok[i + 1]), float(tok[i + 2])); i += 4
    i = tok.index("$Elements") + 1
    m = int(tok[i]); i += 1
    ren, xy, conn, groups = {}, [], [], []
    for _ in range(m):
        typ, ntags = int(tok[i + 1]), int(tok[i + 2])
        phys = int(tok[i + 3]) if ntags > 0 else 0
        nn = {1: 2, 2: 3, 15: 1}[typ]
        vs = [int(t) for t in tok[i + 3 + ntags:i + 3 + ntags + nn]]
        i += 3 + ntags + nn
        if typ == 15:
            continue
        for v in vs:
            if v not in ren:
                ren[v] = len(ren); xy.append(nodes[v])
        if typ == 2:
            conn.append([ren[v] for v in vs]); groups.append(phys)
    return np.array(xy, dtype=np.float64), np.array(conn, dtype=np.int64), np.array(groups)


def refine(xy, conn, groups):
    """GmshMesh::globalRefine(1) of the facade: red refinement, midpoints appended in order of creation (triangles first)."""
    xy = [tuple(p) for p in xy]
    mid = {}

    def midpoint(a, b):
        key = (min(a, b), max(a, b))
        if key not in mid:
            mid[key] = len(xy)
            xy.append((0.5 * (xy[a][0] + xy[b][0]), 0.5 * (xy[a][1] + xy[b][1])))
        return mid[key]
    out, g = [], []
    for e in range(conn.shape[0]):
        v0, v1, v2 = (int(v) for v in conn[e])
        m01, m02, m12 = midpoint(v0, v1), midpoint(v0, v2), midpoint(v1, v2)
        out += [[v0, m01, m02], [m01, v1, m12], [m02, m12, v2], [m01, m12, m02]]
        g += [groups[e]] * 4
    return np.array(xy, dtype=np.float64), np.array(out, dtype=np.int64), np.array(g)


# ---- golden fixtures (tests/golden/um, generated by tests/golden/make_golden_um.py) ---------------------------------------------
def golden_cases():
    import glob
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "um", "*.npz")))


def check_golden(lib, path, solve=None):
    """Drive `lib` (oracle or device binding) through the fixture's problem and compare: integers bit-exact, entries 1e-12 of the
    largest entry, solution 1e-10."""
    g = np.load(path)
    kw = {str(k): float(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    for key in ("with_mass", "shared_space", "cvcf"):
        if key in kw:
            kw[key] = bool(kw[key])
    H = setup(lib, g["xy"], g["conn"], g["sd"], **kw)
    assert H["ndof"] == int(g["ndof"]) and H["ncon"] == int(g["ncon"])
    assert np.array_equal(lib.get_child_offsets(), g["offsets"])
    ptr, dofs = lib.get_dofmap()
    assert np.array_equal(ptr, g["cell_ptr"]) and np.array_equal(dofs, g["cell_dofs"])
    assert np.array_equal(lib.get_constraints(), g["constrained"])
    rp, ci = lib.matrix_get_pattern(H["mat"])
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colidx"])
    x0 = g["x0"].copy()
    assert np.abs(H["x0"] - x0).max() <= 1e-15
    lib.jacobian(H["go"], H["mat"], x0, overwrite=True)
    val = lib.matrix_get_values(H["mat"])
    assert np.abs(val - g["values"]).max() <= 1e-12 * np.abs(g["values"]).max()
    r = lib.residual(H["go"], x0, np.zeros(H["ndof"]))
    assert np.abs(r - g["residual"]).max() <= 1e-12 * np.abs(g["residual"]).max()
    if solve is not None:
        u = x0 - solve(lib, H, r)
        assert np.abs(u - g["solution"]).max() <= 1e-10 * np.abs(g["solution"]).max()
