"""Parity of the CUDA Stokes-Darcy path (csrc/ufe.cu, csrc/direct.cu: SURVEY §8 rows a5 + a12 — multi-child Taylor-Hood subproblem,
ConvectionDiffusionDG on orthonormal Pk triangles, StokesDarcyCouplingOperator with the epsilon finite-difference Jacobian, a direct
solve in place of SuperLU) against the Python oracle (oracle/um_sd_oracle.py), through the C ABI.  Bars (north star): DOF map,
constraints, pattern bit-exact; residual and matrix entries 1e-12 of the largest entry; solutions 1e-10."""
import importlib

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import sd_problems as sdp

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi

RTOL_ENTRY = 1e-12
RTOL_SOLUTION = 1e-10


def _params(**kw):
    return sdp.ini_params(**kw)


CASES = {
    # the shipped configuration on small generated filtration meshes (test/stokesdarcy2.cc + test/topbottomfiltration.ini)
    "sd_6x4": dict(mesh=(6, 4, 3)),
    "sd_12x8": dict(mesh=(12, 8, 11)),
    "sd_intended_operator": dict(mesh=(8, 4, 5), params=dict(quirk=0)),                # the vsize / dsize loop quirk switched off
    "sd_analytic_jacobian": dict(mesh=(6, 4, 7), jac_mode=capi.JAC_ANALYTIC),
    "sd_small_epsilon": dict(mesh=(6, 4, 9), params=dict(epsilon=1e-8)),               # NumericalJacobianCoupling's default epsilon
    "sd_darcy_p2": dict(mesh=(6, 4, 13), darcy_fe=capi.FE_OPB2),
    "sd_darcy_p1": dict(mesh=(6, 4, 14), darcy_fe=capi.FE_OPB1),
    "sd_velocity_dirichlet": dict(mesh=(8, 4, 15), params=dict(stokes_bc=capi.BC_ALL_DIRICHLET)),
    "sd_stokes_only": dict(mesh=(6, 4, 17), darcy=False, params=dict(stokes_bc=capi.BC_ALL_DIRICHLET)),
    "sd_darcy_only": dict(mesh=(6, 4, 19), stokes=False),
    "sd_no_coupling": dict(mesh=(6, 4, 21), coupling=False),
}


def _make(case):
    kw = dict(CASES[case])
    nx, ny, seed = kw.pop("mesh")
    pk = kw.pop("params", None)
    xy, conn, groups = sdp.filtration_mesh(nx, ny, seed=seed)
    dev, ora = pkg.Mdb(0), sdp.oracle()
    Hd = sdp.setup(dev, xy, conn, groups, params=_params(**pk) if pk else None, **kw)
    Ho = sdp.setup(ora, xy, conn, groups, params=_params(**pk) if pk else None, **kw)
    return dev, ora, Hd, Ho


@pytest.fixture(params=sorted(CASES))
def pair(request):
    dev, ora, Hd, Ho = _make(request.param)
    yield dev, ora, Hd, Ho
    dev.close()


def _close(a, b, rtol):
    return np.abs(a - b).max() <= rtol * max(np.abs(b).max(), 1e-300)


def test_numbering_constraints_pattern_bit_exact(pair):
    dev, ora, Hd, Ho = pair
    assert Hd["ndof"] == Ho["ndof"] and Hd["ncon"] == Ho["ncon"]
    assert np.array_equal(dev.get_child_offsets(), ora.get_child_offsets())
    pd, dd = dev.get_dofmap(); po, do = ora.get_dofmap()
    assert np.array_equal(pd, po) and np.array_equal(dd, do)
    assert np.array_equal(dev.get_constraints(), ora.get_constraints())
    rd, cd = dev.matrix_get_pattern(Hd["mat"]); ro, co = ora.matrix_get_pattern(Ho["mat"])
    assert np.array_equal(rd, ro) and np.array_equal(cd, co)


def test_residual_matches_oracle(pair):
    dev, ora, Hd, Ho = pair
    x = Hd["x0"]
    for w in (1.0, -0.37):
        rd = dev.residual(Hd["go"], x, np.zeros(Hd["ndof"]), weight=w)
        ro = ora.residual(Ho["go"], x, np.zeros(Ho["ndof"]), weight=w)
        assert _close(rd, ro, RTOL_ENTRY)
    # additive: r += R(x)
    r0 = np.linspace(-1.0, 1.0, Hd["ndof"])
    rd = dev.residual(Hd["go"], x, r0.copy())
    ro = ora.residual(Ho["go"], x, r0.copy())
    assert _close(rd, ro, RTOL_ENTRY)


def test_jacobian_matches_oracle(pair):
    dev, ora, Hd, Ho = pair
    x = Hd["x0"]
    dev.jacobian(Hd["go"], Hd["mat"], x, overwrite=True)
    ora.jacobian(Ho["go"], Ho["mat"], x, overwrite=True)
    vd, vo = dev.matrix_get_values(Hd["mat"]), ora.matrix_get_values(Ho["mat"])
    assert _close(vd, vo, RTOL_ENTRY)
    # block by block: every child block pair is close relative to ITS largest entry (blocks differ by ten orders of magnitude); the
    # second term is the rounding noise of the finite-difference coupling entries (1e7-sized residuals differenced at epsilon),
    # which also lands in blocks whose exact entries are small
    off = dev.get_child_offsets()
    rowptr, colidx = dev.matrix_get_pattern(Hd["mat"])
    rows = np.repeat(np.arange(Hd["ndof"]), np.diff(rowptr))
    rb, cb = np.searchsorted(off, rows, side="right") - 1, np.searchsorted(off, colidx, side="right") - 1
    for a in range(len(off) - 1):
        for b in range(len(off) - 1):
            sel = (rb == a) & (cb == b)
            if sel.any() and np.abs(vo[sel]).max() > 0:
                assert np.abs(vd[sel] - vo[sel]).max() <= 1e-11 * np.abs(vo[sel]).max() + 1e-13 * np.abs(vo).max(), (a, b)
    # additive with a weight: a += w J; afterwards the constrained rows are unit rows again (jacobianassemblerengine.hh:480-504)
    dev.jacobian(Hd["go"], Hd["mat"], x, weight=-0.5)
    v2 = dev.matrix_get_values(Hd["mat"])
    free = (dev.get_constraints() == 0)[rows]
    assert _close(v2[free], 0.5 * vo[free], 4 * RTOL_ENTRY)
    assert np.array_equal(v2[~free], vo[~free])


def _csr(lib, H):
    rowptr, colidx = lib.matrix_get_pattern(H["mat"])
    return sps.csr_matrix((lib.matrix_get_values(H["mat"]), colidx, rowptr), shape=(H["ndof"], H["ndof"]))


def test_direct_solve_of_the_well_conditioned_stokes_problem():
    """Taylor-Hood with velocity Dirichlet data and the manufactured solution of tests/test_oracle_sd.py: the Newton step from the
    interpolant's perturbation returns the interpolant; banded LU against SciPy's SuperLU to 1e-10."""
    xy, conn, groups = sdp.filtration_mesh(10, 10, seed=5, X=2.0, Y=2.0, y_if=0.0)
    th, da, cp = sdp.ini_params(stokes_bc=capi.BC_ALL_DIRICHLET)
    th.mu = 0.7
    th.f[0], th.f[1] = -2 * 0.7 - 2.0, -2 * 0.7 + 1.0
    dev = pkg.Mdb(0)
    H = sdp.setup(dev, xy, conn, groups, params=(th, da, cp), darcy=False)
    ptr, dofs = dev.get_dofmap()
    xe = np.zeros(H["ndof"])
    for e in range(conn.shape[0]):
        p = xy[conn[e]]
        d = dofs[ptr[e]:ptr[e + 1]]
        nodes = [p[0], 0.5 * (p[0] + p[1]), p[1], 0.5 * (p[0] + p[2]), 0.5 * (p[1] + p[2]), p[2]]
        for i, q in enumerate(nodes):
            xe[d[i]], xe[d[6 + i]] = q[1] ** 2, q[0] ** 2
        for i in range(3):
            xe[d[12 + i]] = 3.0 - 2.0 * p[i][0] + p[i][1]
    con = dev.get_constraints() == 1
    x = xe.copy()
    rng = np.random.default_rng(0)
    x[~con] += rng.uniform(-1, 1, int((~con).sum()))                   # wrong interior values, right Dirichlet values
    # pressure is fixed by nothing but the equations (velocity Dirichlet everywhere -> constant pressure mode): pin it through the start
    dev.jacobian(H["go"], H["mat"], x, overwrite=True)
    r = dev.residual(H["go"], x, np.zeros(H["ndof"]))
    A = _csr(dev, H)
    # remove the constant-pressure null space for the comparison by fixing one pressure coefficient in both solves
    off = dev.get_child_offsets()
    pin = int(off[2])
    A = A.tolil(); A[pin, :] = 0.0; A[pin, pin] = 1.0; A = A.tocsr(); r[pin] = 0.0
    z_ref = spla.spsolve(A.tocsc(), r)
    # the device solves the matrix it holds: overwrite the pinned row there as well (host round trip of the values)
    rowptr, colidx = dev.matrix_get_pattern(H["mat"])
    Ad = sps.csr_matrix((dev.matrix_get_values(H["mat"]), colidx, rowptr), shape=A.shape)
    assert np.abs(Ad[pin, :].toarray()).max() > 0
    dev.set_constraints(np.concatenate([np.nonzero(con)[0], [pin]]))   # a constrained row becomes a unit row in jacobian()
    dev.jacobian(H["go"], H["mat"], x, overwrite=True)
    r2 = dev.residual(H["go"], x, np.zeros(H["ndof"]))
    z = np.zeros(H["ndof"])
    st = dev.solve(H["mat"], r2, z, method=capi.SOLVER_DIRECT, precond=capi.PRECOND_NONE)
    assert st.converged == 1 and st.defect <= 1e-12 * st.defect0
    assert np.abs(z - z_ref).max() <= RTOL_SOLUTION * np.abs(z_ref).max()
    xn = x - z
    vel = np.arange(H["ndof"]) < off[2]
    assert np.abs((xn - xe)[vel]).max() < 1e-10                       # the quadratic velocity is reproduced exactly
    dev.close()


@pytest.mark.parametrize("case", ["sd_6x4", "sd_12x8", "sd_intended_operator"])
def test_stationary_direct_solve_of_the_coupled_problem(case):
    """StationaryLinearProblemSolver(gridoperator, ISTLBackend_SEQ_SuperLU, 1e-15).apply(u) from u = 0 (test/stokesdarcy2.cc:778-790).
    The coupled matrix has entries from 1e-10 (low permeability) to 1e7 (Beavers-Joseph): the bar on the solution is scaled by what
    SciPy's SuperLU itself reproduces (its own forward error estimate from a perturbed right-hand side)."""
    dev, ora, Hd, Ho = _make(case)
    u = np.zeros(Hd["ndof"])
    st = dev.stationary_solve(Hd["go"], Hd["mat"], u, method=capi.SOLVER_DIRECT, precond=capi.PRECOND_NONE)
    assert st.converged == 1
    ora.jacobian(Ho["go"], Ho["mat"], np.zeros(Ho["ndof"]), overwrite=True)
    r = ora.residual(Ho["go"], np.zeros(Ho["ndof"]), np.zeros(Ho["ndof"]))
    A = _csr(ora, Ho).tocsc()
    lu = spla.splu(A)
    z = lu.solve(r)
    z = z + lu.solve(r - A @ z)
    u_ref = -z
    # the defect of the device solution in the ORACLE's system
    assert np.linalg.norm(A @ (-u) - r) <= 1e-10 * np.linalg.norm(r)
    assert st.defect <= 1e-10 * st.defect0
    # forward error: relative to each child block's own scale
    off = dev.get_child_offsets()
    z2 = lu.solve(r * (1.0 + 1e-15 * np.sign(np.sin(np.arange(r.size)))))
    for a in range(len(off) - 1):
        blk = slice(off[a], off[a + 1])
        scale = np.abs(u_ref[blk]).max()
        noise = np.abs(z2[blk] - z[blk]).max()
        assert np.abs(u[blk] - u_ref[blk]).max() <= max(RTOL_SOLUTION * scale, 1e3 * noise), (a, scale, noise)
    # and the residual at the solution vanishes (the problem is affine)
    rs = dev.residual(Hd["go"], u, np.zeros(Hd["ndof"]))
    assert np.linalg.norm(rs) <= 1e-9 * np.linalg.norm(r)
    dev.close()


@pytest.mark.parametrize("path", sdp.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_device_reproduces_golden(path):
    """The committed fixtures of tests/golden/sd (frozen oracle output incl. SciPy's SuperLU solution of the Newton step) through the C ABI."""
    dev = pkg.Mdb(0)

    def solve(H, r):
        z = np.zeros(H["ndof"])
        st = dev.solve(H["mat"], r, z, method=capi.SOLVER_DIRECT, precond=capi.PRECOND_NONE)
        assert st.converged == 1
        return z

    sdp.check_golden(dev, path, solve)
    dev.close()
