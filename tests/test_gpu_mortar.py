"""Parity of the CUDA enriched-coupling (mortar) path against the CPU oracle — SURVEY §8 a11, BASELINE configs[3]
(test/testmortarpoisson.cc: Q1 | Q1 | Q1 mortar space on the interface, MortarPoissonCoupling, FD Jacobians).
Bars as everywhere: DOF map / constraints / pattern bit-exact, entries 1e-12, solutions 1e-10."""
import importlib

import numpy as np
import pytest

import orc
import problems

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi

RTOL_ENTRY = 1e-12
RTOL_SOLUTION = 1e-10


def _stairs(n):
    """A non-planar interface: subdomain 1 = cells right of a staircase in the (x0, x1) plane."""
    idx = np.indices(n[::-1])[::-1]            # idx[d] = cell coordinate along axis d, array in (.., x1, x0) order
    right = idx[0] >= n[0] // 2 + (idx[1] % 3) - 1
    return np.where(right, 2, 1).astype(np.uint8).ravel()


CASES = {
    "m2d_shipped_32": dict(n=(32, 32)),                                    # testmortarpoisson "5 1.0"
    "m2d_ragged": dict(n=(10, 7), gamma_0=3.0),
    "m2d_split_y": dict(n=(6, 12), split_axis=1),
    "m2d_q1q2": dict(n=(8, 8), fe=(capi.FE_Q1, capi.FE_Q2)),
    "m2d_q2q2": dict(n=(6, 4), fe=(capi.FE_Q2, capi.FE_Q2), gamma_0=0.5),
    "m2d_analytic": dict(n=(12, 8), jac_mode=capi.JAC_ANALYTIC),
    "m3d_8": dict(n=(8, 8, 8)),
    "m3d_ragged_split_z": dict(n=(5, 6, 4), split_axis=2, gamma_0=2.0),
    "m3d_q2q1": dict(n=(4, 3, 3), fe=(capi.FE_Q2, capi.FE_Q1)),
    "m2d_stairs": dict(n=(12, 9), sd_bits="stairs"),
    "m3d_stairs": dict(n=(8, 6, 3), sd_bits="stairs"),
    # test/testpowermortarpoisson.cc: PowerCouplingGridFunctionSpace<CouplingGFS,2,VBE>, the operator loops over the blocks
    "p2d_shipped_32": dict(n=(32, 32), ncomp=2),                           # testpowermortarpoisson "5 1.0"
    "p2d_ragged_q1q2": dict(n=(6, 5), ncomp=2, fe=(capi.FE_Q1, capi.FE_Q2), gamma_0=3.0),
    "p2d_analytic_split_y": dict(n=(6, 8), ncomp=2, split_axis=1, jac_mode=capi.JAC_ANALYTIC),
    "p3d_6": dict(n=(6, 5, 4), ncomp=2),
    "p3d_q2q1": dict(n=(4, 3, 3), ncomp=2, fe=(capi.FE_Q2, capi.FE_Q1)),
    "p2d_stairs": dict(n=(12, 9), ncomp=2, sd_bits="stairs"),
}


def _pair(case):
    kw = dict(CASES[case])
    if kw.get("sd_bits") == "stairs":
        kw["sd_bits"] = _stairs(kw["n"])
    dev, ora = pkg.Mdb(0), orc.Oracle()
    return dev, ora, problems.testmortarpoisson(dev, **kw), problems.testmortarpoisson(ora, **kw)


@pytest.fixture(params=sorted(CASES))
def pair(request):
    dev, ora, Pd, Po = _pair(request.param)
    yield dev, ora, Pd, Po
    dev.close(); ora.close()


def _x(ora, Po, kind):
    if kind == "g":
        u = np.zeros(Po.ndof)
        ora.interpolate(Po.sps, Po.gfn, u)
        return u
    return np.random.default_rng(0).uniform(-1, 1, Po.ndof)


def test_dofmap_constraints_pattern_bit_exact(pair):
    dev, ora, Pd, Po = pair
    assert Pd.ndof == Po.ndof and Pd.ncon == Po.ncon
    assert np.array_equal(dev.get_child_offsets(), ora.get_child_offsets())
    pd_, dd = dev.get_dofmap(); po, do = ora.get_dofmap()
    assert np.array_equal(pd_, po) and np.array_equal(dd, do)
    assert np.array_equal(dev.get_constraints(), ora.get_constraints())
    assert dev.nnz[Pd.mat] == ora.nnz[Po.mat]
    rd, cd = dev.matrix_get_pattern(Pd.mat); ro, co = ora.matrix_get_pattern(Po.mat)
    assert np.array_equal(rd, ro) and np.array_equal(cd, co)


@pytest.mark.parametrize("xkind", ["g", "random"])
def test_jacobian_entries(pair, xkind):
    dev, ora, Pd, Po = pair
    x = _x(ora, Po, xkind)
    dev.jacobian(Pd.go, Pd.mat, x, overwrite=True)
    ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    vd, vo = dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat)
    assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()
    dev.jacobian(Pd.go, Pd.mat, x, weight=0.5, overwrite=False)           # additive, weighted (gridoperator.hh:94-97)
    ora.jacobian(Po.go, Po.mat, x, weight=0.5, overwrite=False)
    vd, vo = dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat)
    assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()


@pytest.mark.parametrize("xkind", ["g", "random"])
def test_residual_entries(pair, xkind):
    dev, ora, Pd, Po = pair
    x = _x(ora, Po, xkind)
    rd, ro = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.residual(Pd.go, x, rd); ora.residual(Po.go, x, ro)
    ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    A = ora.csr(Po.mat)
    scale = abs(A) @ np.abs(x) + np.abs(ro)
    scale = np.maximum(scale, np.abs(ro).max() * 1e-3)
    assert np.all(np.abs(rd - ro) <= RTOL_ENTRY * scale)
    rd2, ro2 = rd.copy(), ro.copy()
    dev.residual(Pd.go, x, rd2, weight=0.25); ora.residual(Po.go, x, ro2, weight=0.25)
    assert np.all(np.abs(rd2 - ro2) <= 2 * RTOL_ENTRY * scale)


def test_spmv(pair):
    dev, ora, Pd, Po = pair
    x = _x(ora, Po, "random")
    dev.jacobian(Pd.go, Pd.mat, x, overwrite=True); ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    yd, yo = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.spmv(Pd.mat, x, yd); ora.spmv(Po.mat, x, yo)
    A = ora.csr(Po.mat)
    assert np.all(np.abs(yd - yo) <= 1e-13 * (abs(A) @ np.abs(x) + 1e-300))


def test_stationary_solve_bcgs_ssor(pair):
    """pdesolver.apply(u) of testmortarpoisson.cc:456-463: ISTLBackend_SEQ_BCGS_SSOR (BiCGSTAB + SeqSSOR(3)) on the system with
    the mortar block; same algorithm on both sides — iteration counts agree (+-1), solutions to 1e-10."""
    dev, ora, Pd, Po = pair
    u0 = _x(ora, Po, "g")
    dev.jacobian(Pd.go, Pd.mat, u0, overwrite=True); ora.jacobian(Po.go, Po.mat, u0, overwrite=True)
    rd, ro = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.residual(Pd.go, u0, rd); ora.residual(Po.go, u0, ro)
    zd, zo = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.set_ssor_sweeps(3)
    sd = dev.solve(Pd.mat, rd, zd, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-13, maxit=5000)
    so = ora.solve(Po.mat, ro, zo, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, sweeps=3, reduction=1e-13, maxit=5000)
    if not so.converged:
        # p3d_q2q1: the reference's solver configuration itself stagnates on this synthetic power-space system (defect 3e-2 after
        # 5000 iterations in the oracle) — the device must not claim otherwise; the entries are compared by the other tests
        assert not sd.converged
        return
    assert sd.converged
    # BiCGSTAB on this indefinite system is sensitive to the reduction order of its dot products once the defect nears
    # 1e-13: the counts agree to a few iterations, not exactly
    assert abs(sd.iterations - so.iterations) <= max(2, so.iterations // 6), (sd.iterations, so.iterations)
    diff = zd - zo
    if Po.ncomp > 1:
        # The power operator as shipped (every block adds the element-element terms again) can be singular up to the FD noise:
        # on the 2-D default invocation the exact Jacobian has a null vector (u_left on the interface = 2 lambda_0 = 2 lambda_1 =
        # const; sigma_min = 2.5e-10 with FD entries, 6e-15 analytic).  The solution is then defined up to that vector only, so
        # its component is taken out of the comparison (nothing is removed when the matrix is regular, e.g. in 3-D).
        U, S, Vt = np.linalg.svd(ora.csr(Po.mat).toarray())
        for k in np.nonzero(S < 1e-7 * S[0])[0]:
            diff = diff - Vt[k] * (Vt[k] @ diff)
        A = ora.csr(Po.mat)
        assert np.abs(A @ zd - ro).max() <= 1e-11 * np.abs(ro).max()
    assert np.abs(diff).max() <= RTOL_SOLUTION * np.abs(zo).max()


def test_host_vector_upload_covers_the_mortar_coefficients():
    """mdb_jacobian with x in pinned host memory fetches only the coefficients the FD kernels read: these now include x_c."""
    torch = pytest.importorskip("torch")
    dev, ora, Pd, Po = _pair("m3d_8")
    x = np.random.default_rng(2).uniform(-1, 1, Po.ndof)
    xp = torch.from_numpy(x.copy()).pin_memory()
    dev.jacobian(Pd.go, Pd.mat, xp.numpy(), overwrite=True)
    ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    vd, vo = dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat)
    assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()
    dev.close(); ora.close()


@pytest.mark.parametrize("refinement,scaling,dim,ncomp", [(5, 1.0, 2, 1), (4, 2.5, 2, 1), (3, 1.0, 3, 1), (5, 1.0, 2, 2), (3, 2.0, 3, 2)])
def test_cxx_mortar_driver_matches_oracle(refinement, scaling, dim, ncomp):
    """examples/testmortarpoisson.cc / testpowermortarpoisson.cc (test/test[power]mortarpoisson.cc against include/mdb200/multidomain.hh)
    run as programs; (5, 1.0, 2) is the reference's default invocation: ordering size 1155 (1188 with the power space), BiCGSTAB +
    SSOR(3) to 1e-10."""
    import re
    import subprocess
    import tempfile
    from test_capi import _build_example
    n = (1 << refinement,) * dim
    ora = orc.Oracle()
    Po = problems.testmortarpoisson(ora, n, gamma_0=scaling, ncomp=ncomp)
    u = np.zeros(Po.ndof)
    ora.interpolate(Po.sps, Po.gfn, u)
    ora.jacobian(Po.go, Po.mat, u, overwrite=True)
    r = np.zeros(Po.ndof)
    ora.residual(Po.go, u, r)
    z = np.zeros(Po.ndof)
    so = ora.solve(Po.mat, r, z, reduction=1e-10)
    ora.solve(Po.mat, r, z, reduction=1e-13)
    un = u - z
    with tempfile.TemporaryDirectory() as d:
        exe = _build_example(d, "testmortarpoisson" if ncomp == 1 else "testpowermortarpoisson")
        out = subprocess.run([exe, str(refinement), str(scaling), str(dim)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert int(out.stdout.split()[0]) == Po.ndof
    m = re.search(r"(\d+) dof total, (\d+) dof constrained, (\d+) nonzeroes", out.stdout)
    assert m and int(m.group(1)) == Po.ndof and int(m.group(2)) == Po.ncon and int(m.group(3)) == ora.nnz[Po.mat]
    m = re.search(r"\|r\|_2 = (\S+)", out.stdout)
    assert abs(float(m.group(1)) - np.linalg.norm(r)) <= 1e-11 * np.linalg.norm(r)
    m = re.search(r"solver: (\d+) iterations", out.stdout)
    assert abs(int(m.group(1)) - so.iterations) <= max(2, so.iterations // 6), (m.group(1), so.iterations)
    m = re.search(r"sum\(u\) = (\S+)\s+\|u\|_2 = (\S+)", out.stdout)
    if ncomp > 1 and dim == 2:
        # power space in 2-D: the shipped system is singular up to FD noise (see test_stationary_solve_bcgs_ssor); the 1e-10 solve
        # carries an O(1) multiple of the null vector (FD noise of the right-hand side / sigma_min) that depends on where the
        # Krylov iteration stops — the checksum is not a property of the problem
        return
    assert abs(float(m.group(1)) - un.sum()) <= 1e-7 * abs(un.sum())
    assert abs(float(m.group(2)) - np.linalg.norm(un)) <= 1e-7 * np.linalg.norm(un)


def test_driver_output_files_match_oracle(tmp_path):
    """What the reference drivers write (test/testmortarpoisson.cc:453, :465-486): jacobian.mat by writeMatrixToMatlab ("row col value",
    1-based, BCRS order) and the solution per subdomain as VTK point data — parsed back and compared with the oracle."""
    import re
    import subprocess
    import tempfile
    from test_capi import _build_example
    n = (8, 8)
    ora = orc.Oracle()
    Po = problems.testmortarpoisson(ora, n)
    u = np.zeros(Po.ndof); ora.interpolate(Po.sps, Po.gfn, u)
    ora.jacobian(Po.go, Po.mat, u, overwrite=True)
    r = np.zeros(Po.ndof); ora.residual(Po.go, u, r)
    z = np.zeros(Po.ndof); ora.solve(Po.mat, r, z, reduction=1e-13)
    un = u - z
    with tempfile.TemporaryDirectory() as d:
        exe = _build_example(d, "testmortarpoisson")
        prefix = str(tmp_path) + "/"
        out = subprocess.run([exe, "3", "1.0", "2", prefix], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    trip = np.loadtxt(prefix + "jacobian.mat")
    rp, ci = ora.matrix_get_pattern(Po.mat)
    rows = np.repeat(np.arange(Po.ndof), np.diff(rp))
    assert np.array_equal(trip[:, 0].astype(int) - 1, rows) and np.array_equal(trip[:, 1].astype(int) - 1, ci)
    vo = ora.matrix_get_values(Po.mat)
    assert np.abs(trip[:, 2] - vo).max() <= 1e-12 * np.abs(vo).max()
    off = ora.get_child_offsets()
    for k, side in enumerate(("left", "right")):
        txt = open(prefix + "testmortarpoisson-%s.vtu" % side).read()
        m = re.search(r'NumberOfPoints="(\d+)" NumberOfCells="(\d+)"', txt)
        assert int(m.group(1)) == off[k + 1] - off[k] and int(m.group(2)) == n[0] * n[1] // 2
        vals = np.array(re.search(r'Name="u"[^>]*>\n(.*?)</DataArray>', txt, re.S).group(1).split(), dtype=float)
        # Q1: the points are the subdomain's vertices in ascending index = the child block's DOF order
        assert np.abs(vals - un[off[k]:off[k + 1]]).max() <= 1e-7 * np.abs(un).max()
        pts = np.array(re.search(r'<Points>\n<DataArray[^>]*>\n(.*?)</DataArray>', txt, re.S).group(1).split(), dtype=float).reshape(-1, 3)
        assert (pts[:, 0].min(), pts[:, 0].max()) == ((0.0, 0.5) if k == 0 else (0.5, 1.0))
