"""Parity of the CUDA unstructured-mesh path (csrc/unstructured.cu: triangle meshes, P1 spaces, Poisson / L2 subproblems,
ProportionalFlowCoupling with the reference's finite-difference Jacobian) against the Python oracle (oracle/um_oracle.py), through
the C ABI.  Bars: DOF map, constraints, pattern bit-exact; residual and matrix entries 1e-12 of the largest entry; solutions 1e-10."""
import importlib

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import um_problems as ump

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi

RTOL_ENTRY = 1e-12
RTOL_SOLUTION = 1e-10

CASES = {
    "um_6x5": dict(mesh=(6, 5, 3)),
    "um_17x13_split_y": dict(mesh=(17, 13, 1), axis=1),
    "um_analytic": dict(mesh=(9, 8, 5), jac_mode=capi.JAC_ANALYTIC),
    "um_shared_space": dict(mesh=(8, 7, 2), shared_space=True),
    "um_with_mass": dict(mesh=(7, 6, 4), with_mass=True),
    "um_no_coupling_dirichlet": dict(mesh=(5, 9, 6), coupling=False, bc=capi.BCType(capi.BC_ALL_DIRICHLET)),
    "um_2x1": dict(mesh=(2, 1, 0)),                            # four triangles, two free DOFs per subdomain at most
    "um_40x30": dict(mesh=(40, 30, 7)),
    # ContinuousValueContinuousFlowCoupling (the coupling of test/testpoisson.cc) with the FD Jacobian / exact derivative
    "um_cvcf": dict(mesh=(8, 6, 9), cvcf=True, k=2.0),
    "um_cvcf_split_y_analytic": dict(mesh=(7, 9, 10), axis=1, cvcf=True, k=10.0, jac_mode=capi.JAC_ANALYTIC),
    "um_cvcf_shared_space": dict(mesh=(6, 6, 12), cvcf=True, k=3.0, shared_space=True),
    # ragged markings: an empty subdomain (its child block has no DOF, the coupling no face) and cells that belong to no subdomain
    # (no DOFs there; the faces towards them are interior faces without a condition, residualassemblerfunctors.hh:80-86)
    "um_empty_right_subdomain": dict(mesh=(6, 5, 13), marks="all_left"),
    "um_holes": dict(mesh=(9, 8, 14), marks="holes", cvcf=True),
}


def _make(case):
    kw = dict(CASES[case])
    nx, ny, seed = kw.pop("mesh")
    axis = kw.pop("axis", 0)
    xy, conn = ump.square_mesh(nx, ny, seed=seed)
    sd = ump.halves(xy, conn, axis=axis)
    marks = kw.pop("marks", None)
    if marks == "all_left":
        sd[:] = 1
    elif marks == "holes":
        c = xy[conn].mean(axis=1)
        sd[(np.abs(c[:, 0] - 0.27) < 0.12) & (np.abs(c[:, 1] - 0.6) < 0.15)] = 0       # a hole inside the left subdomain
        sd[(np.abs(c[:, 0] - 0.5) < 0.1) & (np.abs(c[:, 1] - 0.3) < 0.1)] = 0          # and one across the interface
    dev, ora = pkg.Mdb(0), ump.oracle()
    return dev, ora, ump.setup(dev, xy, conn, sd, **kw), ump.setup(ora, xy, conn, sd, **kw)


@pytest.fixture(params=sorted(CASES))
def pair(request):
    dev, ora, Hd, Ho = _make(request.param)
    yield dev, ora, Hd, Ho
    dev.close()


def _close(a, b, rtol):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() <= rtol * scale


def test_numbering_constraints_pattern_bit_exact(pair):
    dev, ora, Hd, Ho = pair
    assert Hd["ndof"] == Ho["ndof"] and Hd["ncon"] == Ho["ncon"]
    assert np.array_equal(dev.get_child_offsets(), ora.get_child_offsets())
    pd, dd = dev.get_dofmap(); po, do = ora.get_dofmap()
    assert np.array_equal(pd, po) and np.array_equal(dd, do)
    assert np.array_equal(dev.get_constraints(), ora.get_constraints())
    rd, cd = dev.matrix_get_pattern(Hd["mat"]); ro, co = ora.matrix_get_pattern(Ho["mat"])
    assert np.array_equal(rd, ro) and np.array_equal(cd, co)
    assert np.array_equal(Hd["x0"], Ho["x0"]) or _close(Hd["x0"], Ho["x0"], 1e-15)     # exp() may differ in the last place


def test_residual_and_jacobian_entries(pair):
    dev, ora, Hd, Ho = pair
    rng = np.random.default_rng(0)
    for x in (Ho["x0"], rng.uniform(-1.0, 1.0, Ho["ndof"])):
        rd = dev.residual(Hd["go"], x, np.zeros(Hd["ndof"]))
        ro = ora.residual(Ho["go"], x, np.zeros(Ho["ndof"]))
        assert _close(rd, ro, RTOL_ENTRY)
        dev.jacobian(Hd["go"], Hd["mat"], x, overwrite=True)
        ora.jacobian(Ho["go"], Ho["mat"], x, overwrite=True)
        assert _close(dev.matrix_get_values(Hd["mat"]), ora.matrix_get_values(Ho["mat"]), RTOL_ENTRY)
    # additive semantics and weights (gridoperator.hh:89-97; one-step weights localassembler.hh:214-255)
    r0 = rng.uniform(-1.0, 1.0, Ho["ndof"])
    rd = dev.residual(Hd["go"], x, r0.copy(), weight=0.3)
    ro = ora.residual(Ho["go"], x, r0.copy(), weight=0.3)
    assert _close(rd, ro, RTOL_ENTRY)
    if "go1" in Hd:                                              # J = 1.0 * J1 + 0.04 * J0 as OneStepGridOperator builds it
        dev.jacobian(Hd["go1"], Hd["mat"], x, weight=1.0, overwrite=True); dev.jacobian(Hd["go"], Hd["mat"], x, weight=0.04)
        ora.jacobian(Ho["go1"], Ho["mat"], x, weight=1.0, overwrite=True); ora.jacobian(Ho["go"], Ho["mat"], x, weight=0.04)
        assert _close(dev.matrix_get_values(Hd["mat"]), ora.matrix_get_values(Ho["mat"]), RTOL_ENTRY)
        rd = dev.residual(Hd["go1"], x, np.zeros(Hd["ndof"])); ro = ora.residual(Ho["go1"], x, np.zeros(Ho["ndof"]))
        assert _close(rd, ro, RTOL_ENTRY)


@pytest.mark.parametrize("precond", ["jacobi", "ssor", "ilu0"])
def test_stationary_solve(pair, precond):
    """StationaryLinearProblemSolver::apply [UPSTREAM (vii)]: J(x0) z = R(x0), x = x0 - z, against SciPy on the oracle's matrix."""
    dev, ora, Hd, Ho = pair
    x0 = Ho["x0"]
    ora.jacobian(Ho["go"], Ho["mat"], x0, overwrite=True)
    ro = ora.residual(Ho["go"], x0, np.zeros(Ho["ndof"]))
    rp, ci = ora.matrix_get_pattern(Ho["mat"])
    A = sps.csr_matrix((ora.matrix_get_values(Ho["mat"]), ci, rp), shape=(Ho["ndof"],) * 2)
    xo = x0 - spla.spsolve(A.tocsc(), ro)
    dev.jacobian(Hd["go"], Hd["mat"], x0, overwrite=True)
    rd = dev.residual(Hd["go"], x0, np.zeros(Hd["ndof"]))
    z = np.zeros(Hd["ndof"])
    pc = dict(jacobi=capi.PRECOND_JACOBI, ssor=capi.PRECOND_SSOR, ilu0=capi.PRECOND_ILU0)[precond]
    st = dev.solve(Hd["mat"], rd, z, method=capi.SOLVER_BICGSTAB, precond=pc, reduction=1e-12, maxit=5000)
    assert st.converged
    assert _close(x0 - z, xo, RTOL_SOLUTION)


def test_rejects_what_this_build_has_no_kernels_for():
    dev = pkg.Mdb(0)
    xy, conn = ump.square_mesh(3, 3)
    dev.mesh_unstructured(xy, conn)
    with pytest.raises(capi.MdbError):
        dev.add_space(capi.FE_Q1, 0)                             # cube spaces on a simplex mesh
    bad = conn.copy(); bad[0, 1] = bad[0, 0]
    with pytest.raises(capi.MdbError):
        dev.mesh_unstructured(xy, bad)                           # degenerate cell
    bad = conn.copy(); bad[0, 0] = xy.shape[0]
    with pytest.raises(capi.MdbError):
        dev.mesh_unstructured(xy, bad)                           # index out of range
    dev.mesh_unstructured(xy, conn)
    with pytest.raises(capi.MdbError):
        dev.mesh_partition((3, 6), (0, 0), (0.0, 0.0), (1.0, 1.0))   # slab partitions are a structured-mesh feature
    dev.mesh_structured((4, 4))
    with pytest.raises(capi.MdbError):
        dev.add_space(capi.FE_P1, 0)
    dev.close()


def test_size_independent_properties_at_scale():
    """768 x 768 squares = 1.18 M triangles, 0.59 M vertices: the residual is affine in x with the assembled matrix as its
    derivative (R(x) - R(0) = J x through mdb_spmv, exact-derivative coupling), constants lie in the kernel of the stiffness
    rows away from the coupling, and the solve drives the residual to zero."""
    nx = 768
    xy, conn = ump.square_mesh(nx, nx, seed=11, shuffle=False)
    perm = np.random.default_rng(5).permutation(conn.shape[0])
    conn = np.ascontiguousarray(conn[perm])
    sd = ump.halves(xy, conn)
    dev = pkg.Mdb(0)
    H = ump.setup(dev, xy, conn, sd, jac_mode=capi.JAC_ANALYTIC)
    n = H["ndof"]
    assert n == (nx + 1) * (nx + 1) + (nx + 1)                    # the interface vertices are in both subdomains
    rng = np.random.default_rng(1)
    x = rng.uniform(-1.0, 1.0, n)
    dev.jacobian(H["go"], H["mat"], x, overwrite=True)
    r0 = dev.residual(H["go"], np.zeros(n), np.zeros(n))
    r1 = dev.residual(H["go"], x, np.zeros(n))
    con = dev.get_constraints() == 1
    assert 0 < np.count_nonzero(con) < 4 * (nx + 1)
    y = dev.spmv(H["mat"], x, np.zeros(n))                       # free rows keep the columns of constrained DOFs (no elimination)
    lhs = np.where(con, 0.0, r1 - r0)
    rhs = np.where(con, 0.0, y)
    assert np.abs(lhs - rhs).max() <= 1e-10 * np.abs(rhs).max()
    assert np.array_equal(y[con], x[con])                        # unit rows
    rp, ci = dev.matrix_get_pattern(H["mat"])
    assert np.all(np.diff(rp) >= 1) and rp[-1] == ci.size
    starts = rp[1:-1]
    d = np.diff(ci)
    mask = np.ones(d.size, dtype=bool); mask[starts - 1] = False
    assert np.all(d[mask] > 0)                                   # ascending columns inside every row (BCRS order)
    z = np.zeros(n)
    st = dev.solve(H["mat"], r1, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_JACOBI, reduction=1e-12, maxit=20000)
    assert st.converged
    r2 = dev.residual(H["go"], x - z, np.zeros(n))
    assert np.abs(r2).max() <= 1e-9 * np.abs(r1).max()
    dev.close()


@pytest.mark.parametrize("path", ump.golden_cases(), ids=lambda p: p.split("/")[-1][:-4])
def test_device_reproduces_golden(path):
    """The committed fixtures of tests/golden/um (frozen oracle output incl. the SciPy solution) through the C ABI."""
    def solve(dev, H, r):
        z = np.zeros(H["ndof"])
        st = dev.solve(H["mat"], r, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-12, maxit=2000)
        assert st.converged
        return z
    dev = pkg.Mdb(0)
    ump.check_golden(dev, path, solve)
    dev.close()


@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_one_step_theta_on_a_triangle_mesh(theta):
    """mdb_onestep_apply (the instationary glue of test/testinstationarypoisson.cc:320-403: OneStepGridOperator over a spatial and
    a temporal grid operator, time dependent Dirichlet data) on the unstructured path, two steps of the one-step-theta scheme
    against the same stage equations formed from the oracle's residuals and Jacobians and solved by SciPy.  Every step starts both
    sides from the SAME vector: the finite-difference coupling Jacobian turns a 1e-14 difference of the linearisation point into
    1e-8 relative noise of its entries, so two free-running trajectories agree only to ~1e-9 after the first step
    (profiles/scripts/dbg_onestep.py) — a property of the reference's algorithm, not of either implementation."""
    xy, conn = ump.square_mesh(9, 8, seed=21)
    sd = ump.halves(xy, conn)
    dev, ora = pkg.Mdb(0), ump.oracle()
    Hd = ump.setup(dev, xy, conn, sd, with_mass=True, cvcf=True, k=2.0)
    Ho = ump.setup(ora, xy, conn, sd, with_mass=True, cvcf=True, k=2.0)
    g = capi.Function(capi.FN_INSTAT_G, 1.5, 3.0)
    n, dt, t = Ho["ndof"], 0.05, 0.1
    ud = np.random.default_rng(3).uniform(-1.0, 1.0, n)
    free = ora.get_constraints() == 0
    rp, ci = ora.matrix_get_pattern(Ho["mat"])
    for step in range(2):
        new = np.zeros(n)
        st = dev.onestep_apply(Hd["go"], Hd["go1"], Hd["mat"], list(Hd["sp"]), [g, g], t, dt, ud, new, scheme=capi.ONESTEP_THETA, theta=theta,
                               method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-13, maxit=2000)
        assert st.converged
        uo, ud = ud, new                                         # uo: the vector this step started from
        # the oracle's stage: x = g(t + dt) on the Dirichlet DOFs, the old values elsewhere; J = J1 + theta dt J0;
        # R = R1(x) - R1(u_old) + dt ((1 - theta) R0(u_old, t) + theta R0(x, t + dt)); u_new = x - J^-1 R
        x = ora.interpolate(list(Ho["sp"]), [g, g], np.zeros(n), time=t + dt)
        x[free] = uo[free]
        ora.jacobian(Ho["go1"], Ho["mat"], x, weight=1.0, overwrite=True)
        ora.jacobian(Ho["go"], Ho["mat"], x, weight=theta * dt)
        R = np.zeros(n)
        ora.set_time(t)
        ora.residual(Ho["go1"], uo, R, weight=-1.0)
        if theta != 1.0:
            ora.residual(Ho["go"], uo, R, weight=(1.0 - theta) * dt)
        ora.set_time(t + dt)
        ora.residual(Ho["go1"], x, R, weight=1.0)
        ora.residual(Ho["go"], x, R, weight=theta * dt)
        A = sps.csr_matrix((ora.matrix_get_values(Ho["mat"]), ci, rp), shape=(n, n))
        assert _close(ud, x - spla.spsolve(A.tocsc(), R), RTOL_SOLUTION)
        t += dt
    dev.close()
