"""Host-side check (no GPU) of the facade's VTK evaluation for the spaces of the Stokes-Darcy driver (include/mdb200/multidomain.hh:
VTKWriter::point_values / corner_index / local_size / opb_corner_values; reference: multidomain/vtk.hh:83-98 addSolutionToVTKWriter as
test/stokesdarcy2.cc:836-876 calls it for the Stokes and the Darcy grid view).  A small C++ harness runs the static functions on the
cell DOF maps of the Stokes-Darcy ORACLE (the same arrays mdb_get_dofmap returns on the device: tests/test_gpu_sd.py checks that
bit for bit) and prints the point values; Python evaluates the same fields independently from the oracle's bases."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import sd_problems as sdp          # noqa: E402
import um_sd_oracle as sdo         # noqa: E402

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "mdb200/multidomain.hh"
#include "../dune-multidomain_b200/csrc/opb_tables.h"
using namespace mdb200;
static std::vector<long long> read_ints(std::FILE* fp) { long long n; if (std::fscanf(fp, "%lld", &n) != 1) std::exit(2); std::vector<long long> v((size_t)n);
  for (size_t i = 0; i < v.size(); ++i) { if (std::fscanf(fp, "%lld", &v[i]) != 1) std::exit(2); }
  return v; }
static std::vector<double> read_doubles(std::FILE* fp) { long long n; if (std::fscanf(fp, "%lld", &n) != 1) std::exit(2); std::vector<double> v((size_t)n);
  for (size_t i = 0; i < v.size(); ++i) { if (std::fscanf(fp, "%lf", &v[i]) != 1) std::exit(2); }
  return v; }
int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "opb") {                 // corner values: long-double Gram-Schmidt against the generated tables
    for (int k = 1; k <= 3; ++k) {
      std::vector<double> v; VTKWriter::opb_corner_values(k, v);
      const int n = (k + 1) * (k + 2) / 2;
      for (int c = 0; c < 3; ++c) for (int i = 0; i < n; ++i) {
        double t = 0.0;
        for (int j = 0; j <= i; ++j) {
          const int* e = k == 1 ? opb1_exp[j] : k == 2 ? opb2_exp[j] : opb3_exp[j];
          const double l = k == 1 ? opb1_L[i][j] : k == 2 ? opb2_L[i][j] : opb3_L[i][j];
          const bool one = c == 0 ? (e[0] == 0 && e[1] == 0) : c == 1 ? e[1] == 0 : e[0] == 0;
          if (one) t += l;
        }
        std::printf("%d %d %d %.17g %.17g\n", k, c, i, v[(size_t)c * n + i], t);
      }
    }
    return 0;
  }
  std::FILE* fp = std::fopen(argv[1], "r");
  std::vector<long long> sdv = read_ints(fp), ptrv = read_ints(fp), dofv = read_ints(fp), cellv = read_ints(fp), cpv = read_ints(fp), npv = read_ints(fp);
  std::vector<double> x = read_doubles(fp);
  const bool darcy = std::string(argv[2]).substr(0, 5) == "darcy";
  std::vector<double> corner_xy, cell_A;
  if (darcy) { corner_xy = read_doubles(fp); cell_A = read_doubles(fp); }
  std::fclose(fp);
  std::vector<uint8_t> sd(sdv.begin(), sdv.end());
  std::vector<int64_t> ptr(ptrv.begin(), ptrv.end()), dofs(dofv.begin(), dofv.end()), cells(cellv.begin(), cellv.end()), cp(cpv.begin(), cpv.end());
  const int which = std::atoi(argv[3]);                 // subdomain of the writer
  if (std::string(argv[2]).substr(0, 2) == "qk") {      // structured mesh: Qk child on subdomain 0, Qk child on subdomain 1 (testpoisson.cc:168-169)
    int k0, k1, dim; if (std::sscanf(argv[2], "qk:%d:%d:%d", &k0, &k1, &dim) != 3) return 2;
    GridFunctionSpace g0(SubDomainGridView(0), k0), g1(SubDomainGridView(1), k1);
    std::vector<const GridFunctionSpace*> leaves; leaves.push_back(&g0); leaves.push_back(&g1);
    std::vector<double> val((size_t)npv[0], 0.0);
    VTKWriter::point_values(leaves, (size_t)which, dim, 1 << dim, cells, sd, ptr, dofs, x.data(), cp, val);
    for (size_t i = 0; i < val.size(); ++i) std::printf("%d %zu %.17g\n", which, i, val[i]);
    return 0;
  }
  // the spaces of test/stokesdarcy2.cc:609-683: [v_0 | v_1 | p] on the Stokes subdomain (0), the head on the Darcy subdomain (1)
  VectorGridFunctionSpace v(SubDomainGridView(0), 2, 2);
  PkGridFunctionSpace p(SubDomainGridView(0), 1);
  CompositeGridFunctionSpace th({&v, &p});
  OPBGridFunctionSpace phi(SubDomainGridView(1), darcy ? std::atoi(argv[2] + 6) : std::atoi(argv[2]));
  std::vector<const GridFunctionSpace*> leaves;
  const GridFunctionSpace* tops[2] = {&th, &phi};
  for (int t = 0; t < 2; ++t) { std::vector<GridFunctionSpace*> l; const_cast<GridFunctionSpace*>(tops[t])->collect(l); leaves.insert(leaves.end(), l.begin(), l.end()); }
  if (darcy) {                                          // the derived fields of the Darcy piece: argv[4..6] = porosity, gravity, density
    std::vector<double> pp((size_t)npv[0], 0.0), v0(pp), v1(pp);
    VTKWriter::darcy_point_fields(leaves, 3, cells, sd, ptr, dofs, x.data(), cp, corner_xy, cell_A, std::atof(argv[4]), std::atof(argv[5]), std::atof(argv[6]), pp, v0, v1);
    for (size_t i = 0; i < pp.size(); ++i) std::printf("%zu %.17g %.17g %.17g\n", i, pp[i], v0[i], v1[i]);
    return 0;
  }
  for (size_t gi = 0; gi < leaves.size(); ++gi) {
    if (leaves[gi]->gv.subdomain != which) continue;
    std::vector<double> val((size_t)npv[0], 0.0);
    VTKWriter::point_values(leaves, gi, 2, 3, cells, sd, ptr, dofs, x.data(), cp, val);
    for (size_t i = 0; i < val.size(); ++i) std::printf("%zu %zu %.17g\n", gi, i, val[i]);
  }
  return 0;
}
"""


@pytest.fixture(scope="module")
def harness():
    if os.environ.get("MDB_VTK_HARNESS"):                # a prebuilt harness, e.g. one compiled with -fsanitize=address,undefined
        return os.environ["MDB_VTK_HARNESS"]
    d = tempfile.mkdtemp()
    src = os.path.join(d, "h.cc")
    open(src, "w").write(HARNESS)
    exe = os.path.join(d, "h")
    libdir = os.path.join(ROOT, "dune-multidomain_b200")
    if not os.path.exists(os.path.join(libdir, "libmdb200.so")):         # the header-only facade still links the C ABI
        subprocess.check_call(["make", "-j8", "-C", os.path.join(libdir, "csrc")])
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", libdir, "-lmdb200", "-Wl,-rpath," + libdir])
    return exe


def test_opb_corner_values_agree_with_the_generated_tables(harness):
    out = subprocess.check_output([harness, "opb"], text=True).split("\n")
    rows = [l.split() for l in out if l]
    assert len(rows) == 3 * (3 + 6 + 10)
    for k, c, i, mine, table in rows:
        assert abs(float(mine) - float(table)) <= 1e-12 * max(1.0, abs(float(table))), (k, c, i, mine, table)
    # ... and with the oracle's evaluation of the basis at the corners
    for k in (1, 2, 3):
        for c, (x, y) in enumerate(((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))):
            val, _ = sdo.opb_eval(k, x, y)
            mine = [float(r[3]) for r in rows if int(r[0]) == k and int(r[1]) == c]
            assert np.abs(np.array(mine) - val).max() <= 1e-12 * np.abs(val).max()


def _write(fp, a, fmt):
    a = np.asarray(a).ravel()
    fp.write("%d\n" % a.size)
    fp.write(" ".join(fmt % t for t in a) + "\n")


@pytest.mark.parametrize("darcy_k", [3, 1])
@pytest.mark.parametrize("which", [0, 1])
def test_point_values_of_the_stokes_darcy_spaces(harness, darcy_k, which):
    xy, conn, groups = sdp.filtration_mesh(6, 4, seed=5)
    ora = sdp.oracle()
    H = sdp.setup(ora, xy, conn, groups, darcy_fe=sdo.FE_OPB1 + darcy_k - 1, params=sdp.ini_params(darcy_k=darcy_k))
    ptr, dofs = ora.get_dofmap()
    sd = np.where(groups > 0, 2, 1)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1.0, 1.0, H["ndof"])
    cells = np.nonzero((sd >> which) & 1)[0]
    used = np.zeros(len(xy), dtype=bool); used[conn[cells].ravel()] = True
    point_of_vertex = np.full(len(xy), -1); point_of_vertex[used] = np.arange(used.sum())
    cp = point_of_vertex[conn[cells]]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "in.txt")
        with open(path, "w") as fp:
            for a in (sd, ptr, dofs, cells, cp, [used.sum()]):
                _write(fp, a, "%d")
            _write(fp, x, "%.17g")
        out = subprocess.check_output([harness, path, str(darcy_k), str(which)], text=True)
    got = {}
    for line in out.split("\n"):
        if line:
            gi, i, v = line.split()
            got.setdefault(int(gi), {})[int(i)] = float(v)
    off = ora.get_child_offsets()
    if which == 0:
        # leaves 0, 1 (velocity components, P2) and 2 (pressure, P1): the coefficient of the vertex DOF; P2 numbers the subdomain's vertices
        # first, P1 numbers them alone — both in ascending vertex order = the order of the output points
        assert sorted(got) == [0, 1, 2]
        npts = int(used.sum())
        for leaf in (0, 1, 2):
            want = x[off[leaf]:off[leaf] + npts]
            mine = np.array([got[leaf][i] for i in range(npts)])
            assert np.array_equal(mine, want), leaf
    else:
        # leaf 3: the modal polynomial of every cell at its corners, averaged over the cells around the vertex
        assert sorted(got) == [3]
        n = sdo.fe_size(sdo.FE_OPB1 + darcy_k - 1)
        acc = np.zeros(int(used.sum())); cnt = np.zeros(int(used.sum()))
        corners = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))
        cell_rank = {int(e): r for r, e in enumerate(cells)}                      # OPB numbers the subdomain's cells x n
        for e in cells:
            coeff = x[off[3] + cell_rank[int(e)] * n: off[3] + (cell_rank[int(e)] + 1) * n]
            for c in range(3):
                val, _ = sdo.opb_eval(darcy_k, *corners[c])
                pt = point_of_vertex[conn[e][c]]
                acc[pt] += coeff @ val; cnt[pt] += 1
        want = acc / cnt
        mine = np.array([got[3][i] for i in range(len(want))])
        assert np.abs(mine - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("n,ks", [((4, 6), (1, 2)), ((4, 4), (2, 1)), ((3, 2, 4), (1, 1)), ((2, 2, 2), (2, 2))])
@pytest.mark.parametrize("which", [0, 1])
def test_point_values_of_qk_spaces_are_the_corner_coefficients(harness, n, ks, which):
    """The conforming output of the structured drivers (testpoisson: Q1 | Q2): the value at a vertex is the coefficient of the shape
    function at that corner, found through the cell DOF map with the block offsets of the children present on the cell — checked against
    the C++ oracle's DOF map and an independent lattice computation (Appendix B numbering: child blocks lexicographic over the lattice)."""
    import orc
    import problems
    capi = orc.capi
    o = orc.Oracle()
    dim = len(n)
    fe = tuple({1: capi.FE_Q1, 2: capi.FE_Q2}[k] for k in ks)
    P = problems.testpoisson(o, n, fe=fe, split_axis=dim - 1)
    ptr, dofs = o.get_dofmap()
    sd = o.get_subdomains().astype(np.int64)
    x = np.random.default_rng(3).uniform(-1, 1, P.ndof)
    ncells = int(np.prod(n))
    cells = np.nonzero((sd >> which) & 1)[0]
    nv = [m + 1 for m in n]

    def vertex(e, c):                                   # Yasp: cells and vertices lexicographic, corner bit d = offset along axis d
        idx, stride, rem = 0, 1, int(e)
        for d in range(dim):
            idx += (rem % n[d] + ((c >> d) & 1)) * stride
            rem //= n[d]; stride *= nv[d]
        return idx
    used = np.zeros(int(np.prod(nv)), dtype=bool)
    for e in cells:
        for c in range(1 << dim):
            used[vertex(e, c)] = True
    pov = np.full(used.size, -1); pov[used] = np.arange(used.sum())
    cp = np.array([[pov[vertex(e, c)] for c in range(1 << dim)] for e in cells])
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "in.txt")
        with open(path, "w") as fp:
            for a in (sd, ptr, dofs, cells, cp, [used.sum()]):
                _write(fp, a, "%d")
            _write(fp, x, "%.17g")
        out = subprocess.check_output([harness, path, "qk:%d:%d:%d" % (ks[0], ks[1], dim), str(which)], text=True)
    mine = np.array([float(l.split()[2]) for l in out.split("\n") if l])
    # independent: a Qk block numbers its vertex coefficients first, by the rank of the vertex among the subdomain's vertices in
    # lexicographic order (Appendix B for Q1; [UPSTREAM] (iii) entity groups for Q2: vertices, then the higher-dimensional entities) —
    # and the output points are the subdomain's vertices in the same order
    off = o.get_child_offsets()[which]
    want = x[off:off + int(used.sum())]
    assert ncells == len(sd) and np.array_equal(mine, want)


@pytest.mark.parametrize("darcy_k", [3, 2])
def test_derived_darcy_fields(harness, darcy_k):
    """VTKPressureFromPotential and VTKDarcyFlowFromPotential (test/stokesdarcy2.cc:443-560) from the orthonormal coefficients: p = (phi - y)
    g rho and v = -A grad(phi) / porosity per cell corner (A by physical group), averaged per vertex — against the oracle's basis, its
    reference gradients and its cell geometry (J^-T)."""
    xy, conn, groups = sdp.filtration_mesh(6, 4, seed=11)
    ora = sdp.oracle()
    H = sdp.setup(ora, xy, conn, groups, darcy_fe=sdo.FE_OPB1 + darcy_k - 1, params=sdp.ini_params(darcy_k=darcy_k))
    ptr, dofs = ora.get_dofmap()
    sd = np.where(groups > 0, 2, 1)
    x = np.random.default_rng(9).uniform(-1.0, 1.0, H["ndof"])
    cells = np.nonzero((sd >> 1) & 1)[0]
    used = np.zeros(len(xy), dtype=bool); used[conn[cells].ravel()] = True
    pov = np.full(len(xy), -1); pov[used] = np.arange(used.sum())
    cp = pov[conn[cells]]
    kabs, kabs_low, porosity, gravity, density = 1.16e-5, 1.16e-10, 0.01, 9.81, 1e3
    cell_A = np.where(groups[cells] == 1, kabs, kabs_low)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "in.txt")
        with open(path, "w") as fp:
            for a in (sd, ptr, dofs, cells, cp, [used.sum()]):
                _write(fp, a, "%d")
            for a in (x, xy[conn[cells]], cell_A):
                _write(fp, a, "%.17g")
        out = subprocess.check_output([harness, path, "darcy:%d" % darcy_k, "1", repr(porosity), repr(gravity), repr(density)], text=True)
    got = np.array([[float(t) for t in l.split()[1:]] for l in out.split("\n") if l])
    off = ora.get_child_offsets()
    n = sdo.fe_size(sdo.FE_OPB1 + darcy_k - 1)
    npts = int(used.sum())
    acc = np.zeros((npts, 3)); cnt = np.zeros(npts)
    corners = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))
    for r, e in enumerate(cells):
        coeff = x[off[3] + r * n: off[3] + (r + 1) * n]
        _, jinvt, _ = ora._geo(int(e))                   # J^-T of the affine cell map
        for c in range(3):
            val, grad = sdo.opb_eval(darcy_k, *corners[c])
            g = jinvt @ (coeff @ grad)
            pt = pov[conn[e][c]]
            acc[pt, 0] += (coeff @ val - xy[conn[e][c]][1]) * gravity * density
            acc[pt, 1:] += -cell_A[r] * g / porosity
            cnt[pt] += 1
    want = acc / cnt[:, None]
    for col in range(3):
        assert np.abs(got[:, col] - want[:, col]).max() <= 1e-11 * np.abs(want[:, col]).max(), col
