"""Parity of the CUDA QkDG / ConvectionDiffusionDG path against the CPU oracle — BASELINE configs[0] as the reference ships it
(test/testpoisson-dirichlet-neumann.cc:883-977: DG-Q2 on both subdomains, ConvectionDiffusionDG(SIPG, weightsOn, 2.0),
ContinuousValueContinuousFlowCoupling(4, 10) between them, BiCGSTAB + SSOR(3)); the first operator to use the alpha_skeleton and
alpha_boundary-on-interface dispatch of residualassemblerfunctors.hh:58-120.  Bars as everywhere: DOF map / constraints / pattern
bit-exact, entries 1e-12, solutions 1e-10."""
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
    idx = np.indices(n[::-1])[::-1]
    right = idx[0] >= n[0] // 2 + (idx[1] % 3) - 1
    return np.where(right, 2, 1).astype(np.uint8).ravel()


CASES = {
    "dg2d_q2_shipped_16": dict(n=(16, 16)),                                 # the shipped configuration at mesh.refine = 4
    "dg2d_q2_ragged": dict(n=(7, 5), scaling=3.0, alpha=4.0),
    "dg2d_q1": dict(n=(10, 6), k=1),
    "dg2d_q1_nipg_weights_off": dict(n=(6, 6), k=1, method=capi.DG_METHOD_NIPG, weights=capi.DG_WEIGHTS_OFF),
    "dg2d_q2_uncoupled": dict(n=(6, 4), coupled=False),
    "dg2d_q2_all_dirichlet_poly": dict(n=(5, 4), bc=capi.BCType(capi.BC_ALL_DIRICHLET), g=capi.Function(capi.FN_POLY2, 0.3, 1.0, -2.0, 0.0, 1.5),
                                       f=capi.Function(capi.FN_CONST, -6.0)),
    "dg2d_q2_stairs": dict(n=(12, 9), sd_bits="stairs"),
    "dg3d_q1": dict(n=(6, 5, 4), k=1),
    "dg3d_q2": dict(n=(4, 3, 3)),
    "dg3d_q2_stairs": dict(n=(6, 6, 2), sd_bits="stairs"),
}


def _pair(case):
    kw = dict(CASES[case])
    if kw.get("sd_bits") == "stairs":
        kw["sd_bits"] = _stairs(kw["n"])
    dev, ora = pkg.Mdb(0), orc.Oracle()
    return dev, ora, problems.dirichlet_neumann_dg(dev, **kw), problems.dirichlet_neumann_dg(ora, **kw)


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
    assert Pd.ndof == Po.ndof and Pd.ncon == Po.ncon == 0
    assert np.array_equal(dev.get_child_offsets(), ora.get_child_offsets())
    pd_, dd = dev.get_dofmap(); po, do = ora.get_dofmap()
    assert np.array_equal(pd_, po) and np.array_equal(dd, do)
    assert np.array_equal(dev.get_constraints(), ora.get_constraints())
    assert dev.nnz[Pd.mat] == ora.nnz[Po.mat]
    rd, cd = dev.matrix_get_pattern(Pd.mat); ro, co = ora.matrix_get_pattern(Po.mat)
    assert np.array_equal(rd, ro) and np.array_equal(cd, co)


def test_interpolation(pair):
    dev, ora, Pd, Po = pair
    ud, uo = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.interpolate(Pd.sps, Pd.gfn, ud); ora.interpolate(Po.sps, Po.gfn, uo)
    assert np.abs(ud - uo).max() <= 1e-15 * np.abs(uo).max()


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
    """StationaryLinearProblemSolver with ISTLBackend_SEQ_BCGS_SSOR(5000) (test/testpoisson-dirichlet-neumann.cc:626-640): start from
    u = 0 (a DG space has no Dirichlet extension to interpolate), z = J^-1 R(u), u <- u - z."""
    dev, ora, Pd, Po = pair
    u0 = np.zeros(Po.ndof)
    dev.jacobian(Pd.go, Pd.mat, u0, overwrite=True); ora.jacobian(Po.go, Po.mat, u0, overwrite=True)
    rd, ro = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.residual(Pd.go, u0, rd); ora.residual(Po.go, u0, ro)
    zd, zo = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.set_ssor_sweeps(3)
    sd = dev.solve(Pd.mat, rd, zd, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-13, maxit=5000)
    so = ora.solve(Po.mat, ro, zo, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, sweeps=3, reduction=1e-13, maxit=5000)
    assert sd.converged and so.converged
    assert abs(sd.iterations - so.iterations) <= max(2, so.iterations // 6), (sd.iterations, so.iterations)
    assert np.abs(zd - zo).max() <= RTOL_SOLUTION * np.abs(zo).max()


def test_jacobi_cg_on_the_symmetric_uncoupled_system():
    """SIPG without the (unsymmetric) coupling operator is symmetric positive definite: CG + Jacobi, the device fast path."""
    dev, ora, Pd, Po = _pair("dg2d_q2_uncoupled")
    u0 = np.zeros(Po.ndof)
    dev.jacobian(Pd.go, Pd.mat, u0, overwrite=True); ora.jacobian(Po.go, Po.mat, u0, overwrite=True)
    rd, ro = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.residual(Pd.go, u0, rd); ora.residual(Po.go, u0, ro)
    zd, zo = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    sd = dev.solve(Pd.mat, rd, zd, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-13, maxit=5000)
    so = ora.solve(Po.mat, ro, zo, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-13, maxit=5000)
    assert sd.converged and so.converged and abs(sd.iterations - so.iterations) <= 2
    assert np.abs(zd - zo).max() <= RTOL_SOLUTION * np.abs(zo).max()
    dev.close(); ora.close()


def test_host_vector_upload_with_dg_children():
    """mdb_jacobian with x in pinned host memory gathers only the coefficients of the interface cells (FD coupling Jacobian)."""
    torch = pytest.importorskip("torch")
    dev, ora, Pd, Po = _pair("dg3d_q1")
    x = np.random.default_rng(2).uniform(-1, 1, Po.ndof)
    xp = torch.from_numpy(x.copy()).pin_memory()
    dev.jacobian(Pd.go, Pd.mat, xp.numpy(), overwrite=True)
    ora.jacobian(Po.go, Po.mat, x, overwrite=True)
    vd, vo = dev.matrix_get_values(Pd.mat), ora.matrix_get_values(Po.mat)
    assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()
    dev.close(); ora.close()


def test_shipped_size_properties():
    """mesh.refine = 9 as in poisson-dirichlet-neumann.ini: 512^2 cells, 2 359 296 DOFs, at most 45 entries per row (ini:5-7); SIPG
    rows sum to zero away from Dirichlet faces and the interface (constants are in the kernel of the bilinear form there); the
    solve converges and u - z satisfies R(u - z) = 0 (the operator is affine)."""
    dev = pkg.Mdb(0)
    P = problems.dirichlet_neumann_dg(dev, (512, 512))
    assert P.ndof == 2359296 and P.ncon == 0
    rp, ci = dev.matrix_get_pattern(P.mat)
    assert np.diff(rp).max() == 45 and dev.nnz[P.mat] == 9 * 9 * (5 * 512 * 512 - 4 * 512)
    u0 = np.zeros(P.ndof)
    dev.jacobian(P.go, P.mat, u0, overwrite=True)
    ones, y = np.ones(P.ndof), np.zeros(P.ndof)
    dev.spmv(P.mat, ones, y)
    cell = np.arange(P.ndof) // 9
    half = 512 * 256
    cx = np.where(cell < half, cell % 256, 256 + (cell - half) % 256); cy = np.where(cell < half, cell // 256, (cell - half) // 256)
    inner = (cx > 0) & (cx < 511) & (cy > 0) & (cy < 511) & (cx != 255) & (cx != 256)
    assert np.abs(y[inner]).max() < 1e-9
    r = np.zeros(P.ndof); dev.residual(P.go, u0, r)
    z = np.zeros(P.ndof)
    st = dev.solve(P.mat, r, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_JACOBI, reduction=1e-10, maxit=20000)
    assert st.converged
    r2 = np.zeros(P.ndof); dev.residual(P.go, u0 - z, r2)
    assert np.linalg.norm(r2) <= 1e-8 * np.linalg.norm(r)
    dev.close()
