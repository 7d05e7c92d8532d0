"""the Gmsh 2.x reader, uniform refinement and physical-group -> subdomain marking of the C++ facade
(include/mdb200/multidomain.hh: GmshReader, GmshMesh::globalRefine; reference test/stokesdarcy2.cc:575-604), through the example
driver examples/testgmshpoisson.cc.  CPU part: the reader's output sizes against an independent Python reading of the same file
(the driver stops at the context: no CPU fallback).  GPU part: the driver's DOF counts and solution checksums against the oracle."""
import importlib
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import um_problems as ump

pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi
ROOT = ump.ROOT


def _build(outdir):
    exe = os.path.join(outdir, "testgmshpoisson")
    libdir = os.path.dirname(pkg.library_path())
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "testgmshpoisson.cc"), "-o", exe, "-L", libdir, "-lmdb200", "-Wl,-rpath," + libdir])
    return exe


def _fixture(path, nx=7, ny=5, seed=2, **kw):
    xy, conn = ump.square_mesh(nx, ny, seed=seed)
    groups = np.where(xy[conn].mean(axis=1)[:, 0] > 0.5, 3, 0)          # physical group 3 -> subdomain 1, group 0 -> subdomain 0
    ump.write_gmsh(path, xy, conn, groups, **kw)
    return xy, conn, groups


def test_reader_and_refinement_sizes_on_cpu():
    import torch
    with tempfile.TemporaryDirectory() as d:
        exe = _build(d)
        msh = os.path.join(d, "square.msh")
        _fixture(msh, node_id_offset=5, extra_nodes=3)
        xy, conn, groups = ump.read_gmsh(msh)
        assert xy.shape[0] == 8 * 6 and conn.shape[0] == 70               # the unused nodes and the point element are dropped
        for r in (0, 2):
            p = subprocess.run([exe, msh, str(r)], capture_output=True, text=True)
            rx, rc, rg = xy, conn, groups
            for _ in range(r):
                rx, rc, rg = ump.refine(rx, rc, rg)
            assert "number of real vertices = %d" % xy.shape[0] in p.stdout and "number of elements = 70" in p.stdout
            assert "%d vertices, %d elements after %d refinements" % (rx.shape[0], rc.shape[0], r) in p.stdout
            if not torch.cuda.is_available():
                assert p.returncode != 0 and "mdb_create" in p.stderr        # no CPU fallback
        bad = os.path.join(d, "bad.msh")
        open(bad, "w").write(open(msh).read().replace(" 2 3 ", " 3 3 ", 1))  # a quadrilateral record
        p = subprocess.run([exe, bad], capture_output=True, text=True)
        assert p.returncode != 0 and "not supported" in p.stderr
        p = subprocess.run([exe, os.path.join(d, "missing.msh")], capture_output=True, text=True)
        assert p.returncode != 0 and "cannot open" in p.stderr


def test_reader_and_refinement_entry_by_entry_on_cpu():
    """GmshReader + GmshMesh::globalRefine against the independent Python reading (um_problems.read_gmsh / refine): vertex numbering by
    first appearance in $Elements (boundary lines come first in this fixture, so the numbering differs from the node order), element
    order, physical groups, boundary segments, and the refinement's vertex / child order — all bit-exact."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "gmsh_dump")
        libdir = os.path.dirname(pkg.library_path())
        subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "gmsh_dump.cc"),
                               "-o", exe, "-L", libdir, "-lmdb200", "-Wl,-rpath," + libdir])
        msh = os.path.join(d, "square.msh")
        _fixture(msh, nx=5, ny=4, seed=9, node_id_offset=3, extra_nodes=2)
        xy, conn, groups = ump.read_gmsh(msh)
        assert not np.array_equal(xy, ump.square_mesh(5, 4, seed=9)[0])            # the reader renumbers
        for r in (0, 1, 2):
            tok = subprocess.run([exe, msh, str(r)], capture_output=True, text=True, check=True).stdout.split()
            nv, ne, nb = int(tok[0]), int(tok[1]), int(tok[2])
            rx, rc, rg = xy, conn, groups
            for _ in range(r):
                rx, rc, rg = ump.refine(rx, rc, rg)
            assert (nv, ne, nb) == (rx.shape[0], rc.shape[0], 2 * (5 + 4) * 2 ** r)
            v = np.array(tok[3:3 + 2 * nv], dtype=np.float64).reshape(nv, 2)
            e = np.array(tok[3 + 2 * nv:3 + 2 * nv + 4 * ne], dtype=np.int64).reshape(ne, 4)
            b = np.array(tok[3 + 2 * nv + 4 * ne:], dtype=np.int64).reshape(nb, 3)
            assert np.array_equal(v, rx) and np.array_equal(e[:, :3], rc) and np.array_equal(e[:, 3], rg)
            assert np.all(b[:, 2] == 7)
            on_boundary = lambda p: (np.minimum(np.abs(p[:, 0]), np.abs(p[:, 0] - 1)) == 0) | (np.minimum(np.abs(p[:, 1]), np.abs(p[:, 1] - 1)) == 0)
            assert on_boundary(v[b[:, 0]]).all() and on_boundary(v[b[:, 1]]).all()


REF_MSH = "/root/reference/test"


@pytest.mark.skipif(not os.path.isdir(REF_MSH), reason="the reference's mesh fixtures exist in the build container only")
def test_reader_on_the_reference_fixtures():
    """The C++ reader on the reference's own test/*.msh files (Gmsh 2.1 ASCII, the input of stokesdarcy2): sizes, cells per physical
    group and checksums of coordinates / connectivity in the reader's numbering against tests/golden/gmsh_fixture_sizes.json (written
    by tests/golden/make_golden_gmsh.py with the independent Python reader).  Runs where /root/reference exists; never on the GPU box."""
    import json
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "gmsh_fixture_sizes.json")))
    assert gold["stokesdarcy2-fine.msh"]["triangles"] == 4805
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "gmsh_dump")
        libdir = os.path.dirname(pkg.library_path())
        subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "gmsh_dump.cc"),
                               "-o", exe, "-L", libdir, "-lmdb200", "-Wl,-rpath," + libdir])
        for name, g in sorted(gold.items()):
            tok = subprocess.run([exe, os.path.join(REF_MSH, name), "0"], capture_output=True, text=True, check=True).stdout.split()
            nv, ne, nb = int(tok[0]), int(tok[1]), int(tok[2])
            assert (nv, ne) == (g["vertices"], g["triangles"]), name
            v = np.array(tok[3:3 + 2 * nv], dtype=np.float64).reshape(nv, 2)
            e = np.array(tok[3 + 2 * nv:3 + 2 * nv + 4 * ne], dtype=np.int64).reshape(ne, 4)
            grp, cnt = np.unique(e[:, 3], return_counts=True)
            assert {str(int(a)): int(b) for a, b in zip(grp, cnt)} == g["groups"], name
            assert v[:, 0].sum() == g["sum_x"] and v[:, 1].sum() == g["sum_y"], name
            assert int(e[:, :3].sum()) == g["sum_conn"] and int((e[:, :3] * np.array([1, 3, 7])).sum()) == g["weighted_conn"], name
            p = v[e[:, :3]]
            det = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
            assert np.all(det != 0.0), name                                         # no degenerate cell: mdb_mesh_unstructured would take it


@pytest.mark.gpu
@pytest.mark.parametrize("refinements,solver", [(0, "ssor"), (1, "ssor"), (2, "jac")])
def test_driver_against_oracle(refinements, solver):
    with tempfile.TemporaryDirectory() as d:
        exe = _build(d)
        msh = os.path.join(d, "square.msh")
        _fixture(msh)
        out = subprocess.run([exe, msh, str(refinements), "2.0", solver, d + "/"], capture_output=True, text=True, check=True).stdout
        xy, conn, groups = ump.read_gmsh(msh)
        for _ in range(refinements):
            xy, conn, groups = ump.refine(xy, conn, groups)
        o = ump.oracle()
        sd = np.where(groups > 0, 2, 1).astype(np.uint8)
        H = ump.setup(o, xy, conn, sd, k=2.0, orders=(2, 2), f=capi.Function(capi.FN_CONST, 1.0), bc=capi.BCType(capi.BC_ALL_DIRICHLET))
        x0 = np.zeros(H["ndof"])
        o.jacobian(H["go"], H["mat"], x0, overwrite=True)
        r = o.residual(H["go"], x0, np.zeros(H["ndof"]))
        rp, ci = o.matrix_get_pattern(H["mat"])
        u = x0 - spla.spsolve(sps.csr_matrix((o.matrix_get_values(H["mat"]), ci, rp), shape=(H["ndof"],) * 2).tocsc(), r)
        m = re.search(r"(\d+) dof total, (\d+) dof constrained", out)
        assert (int(m.group(1)), int(m.group(2))) == (H["ndof"], H["ncon"])
        m = re.search(r"sum\(u\) = (\S+)\s+\|u\|_2 = (\S+)", out)
        assert abs(float(m.group(1)) - u.sum()) <= 1e-9 * abs(u.sum())
        assert abs(float(m.group(2)) - np.linalg.norm(u)) <= 1e-9 * np.linalg.norm(u)
        # VTK output per subdomain (vtk.hh:83-98): triangles of the subdomain, its vertices, the solution as point data
        off = o.get_child_offsets()
        for side, sdi in (("left", 0), ("right", 1)):
            txt = open(os.path.join(d, "testgmshpoisson-%s.vtu" % side)).read()
            cells = np.nonzero((sd >> sdi) & 1)[0]
            verts = np.unique(conn[cells])
            assert 'NumberOfPoints="%d" NumberOfCells="%d"' % (verts.size, cells.size) in txt

            def block(name_attr):
                body = txt.split(name_attr, 1)[1].split(">", 1)[1].split("</DataArray>", 1)[0]
                return np.array(body.split(), dtype=np.float64)
            vals = block('Name="u"')
            pts = block('NumberOfComponents="3"').reshape(-1, 3)
            con = block('Name="connectivity"').astype(np.int64).reshape(-1, 3)
            assert np.array_equal(block('Name="types"'), np.full(cells.size, 5.0))
            assert np.abs(pts[:, :2] - xy[verts]).max() == 0.0 and np.array_equal(verts[con], conn[cells])
            assert np.abs(vals - u[off[sdi]:off[sdi + 1]]).max() <= 1e-9 * np.abs(u).max()     # DOF order = subdomain vertex order
