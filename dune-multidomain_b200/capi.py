"""ctypes binding of the C ABI declared in include/mdb200.h.

This is plumbing, not the product: the product is libmdb200.so (hand-written CUDA for sm_100a behind a C ABI).
The binding is generic over a symbol prefix so that the test-suite can drive a second library exporting the same
call shapes (the CPU oracle, prefix ``orc_``) with the very same problem-setup code; the product itself only
ever loads ``libmdb200.so`` and fails loudly when it is missing (there is no CPU fallback).
"""
import ctypes as C
import os

import numpy as np

# ---- enums (mirror include/mdb200.h) -----------------------------------------------------------
FE_Q1, FE_Q2 = 1, 2
FE_DGQ1, FE_DGQ2 = 11, 12
FE_P1 = 21
FE_P2 = 22
FE_OPB1, FE_OPB2, FE_OPB3 = 31, 32, 33
ETYPE_TRIANGLE = 2
COND_EQUALITY, COND_SUBSET = 0, 1
FN_ZERO, FN_CONST, FN_GAUSS, FN_BOX, FN_TESTPOISSON_J, FN_INSTAT_G, FN_POLY2, FN_DN_SOURCE, FN_DARCY_G, FN_DARCY_J, FN_PRESSURE_DROP = range(11)
BC_ALL_DIRICHLET, BC_ALL_NEUMANN, BC_TESTPOISSON, BC_DARCY_RIGHT = range(4)
OP_POISSON, OP_L2, OP_CONVDIFF_DG, OP_TAYLORHOOD_STOKES, OP_DARCY_DG = 1, 2, 3, 4, 5
OP_CVCF_COUPLING, OP_PROPFLOW_COUPLING = 101, 102
OP_STOKESDARCY_COUPLING = 104
OP_MORTAR_POISSON_COUPLING = 201
OP_ND_COUPLING = 103
ND_MODE_NEUMANN, ND_MODE_DIRICHLET, ND_MODE_BOTH = 0, 1, 2
DG_METHOD_NIPG, DG_METHOD_SIPG = 0, 1
DG_WEIGHTS_OFF, DG_WEIGHTS_ON = 0, 1
JAC_FD_BITMATCH, JAC_ANALYTIC = 0, 1
SOLVER_CG, SOLVER_BICGSTAB, SOLVER_DIRECT = 0, 1, 2
PRECOND_NONE, PRECOND_JACOBI, PRECOND_ILU0, PRECOND_SSOR, PRECOND_MG = 0, 1, 2, 3, 4
ONESTEP_IMPLICIT_EULER, ONESTEP_THETA, ONESTEP_ALEXANDER2 = 0, 1, 2
PRECOND_SSOR_ORACLE = 100  # the oracle's own id for ISTL SeqSSOR (tests/orc.py maps PRECOND_SSOR to it)


class Function(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 6)]

    def __init__(self, kind=FN_ZERO, *p):
        super().__init__()
        self.kind = kind
        for i, v in enumerate(p):
            self.p[i] = v


class BCType(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double * 4)]

    def __init__(self, kind=BC_ALL_DIRICHLET, *p):
        super().__init__()
        self.kind = kind
        for i, v in enumerate(p):
            self.p[i] = v


class PoissonParams(C.Structure):
    _fields_ = [("f", Function), ("bc", BCType), ("j", Function), ("quad_order", C.c_int32), ("reserved", C.c_int32)]


class L2Params(C.Structure):
    _fields_ = [("quad_order", C.c_int32), ("reserved", C.c_int32)]


class CVCFParams(C.Structure):
    _fields_ = [("quad_order", C.c_int32), ("reserved", C.c_int32), ("intensity", C.c_double)]


class PropFlowParams(C.Structure):
    _fields_ = [("intensity", C.c_double)]


class NDCouplingParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("method", C.c_int32), ("weights", C.c_int32), ("intorderadd", C.c_int32), ("alpha", C.c_double)]


class ConvDiffDGParams(C.Structure):
    _fields_ = [("f", Function), ("bc", BCType), ("g", Function), ("j", Function), ("method", C.c_int32), ("weights", C.c_int32),
                ("intorderadd", C.c_int32), ("reserved", C.c_int32), ("alpha", C.c_double)]


class TaylorHoodParams(C.Structure):
    """mdb_taylorhood_params: [UPSTREAM] TaylorHoodNavierStokes<NavierStokesParameters>(params, stokes_q), test/stokesdarcy2.cc:226-251, :719-720."""
    _fields_ = [("mu", C.c_double), ("rho", C.c_double), ("f", C.c_double * 2), ("bc", BCType), ("j", Function), ("quad_order", C.c_int32),
                ("full_tensor", C.c_int32)]


class DarcyParams(C.Structure):
    """mdb_darcy_params: [UPSTREAM] ConvectionDiffusionDG<DarcyParameters, OPB>(SIPG, weightsOn, alpha), test/stokesdarcy2.cc:254-407, :733-739."""
    _fields_ = [("kabs", C.c_double), ("kabs_low", C.c_double), ("bc", BCType), ("g", Function), ("j", Function), ("high_group", C.c_int32),
                ("method", C.c_int32), ("weights", C.c_int32), ("intorderadd", C.c_int32), ("alpha", C.c_double)]


class StokesDarcyParams(C.Structure):
    """mdb_stokesdarcy_params: CouplingParameters of test/stokesdarcycouplingoperator.hh:11-77 (+ the vsize/dsize loop quirk switch)."""
    _fields_ = [("gravity", C.c_double), ("alpha", C.c_double), ("viscosity", C.c_double), ("porosity", C.c_double), ("gamma", C.c_double),
                ("density", C.c_double), ("kabs", C.c_double), ("epsilon", C.c_double), ("quirk", C.c_int32), ("reserved", C.c_int32)]


class MortarParams(C.Structure):
    _fields_ = [("gamma_0", C.c_double)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("defect0", C.c_double),
                ("defect", C.c_double), ("seconds", C.c_double)]


def gauss(scale=1.0, centre=(0.5, 0.5, 0.5)):
    """G of test/testpoisson.cc:92-99."""
    return Function(FN_GAUSS, scale, *centre)


class MdbError(RuntimeError):
    pass


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _vecptr(v):
    """Host numpy array, raw integer device pointer, or an object with data_ptr() (torch tensor)."""
    if v is None:
        return C.c_void_p(0)
    if isinstance(v, np.ndarray):
        assert v.dtype == np.float64 and v.flags["C_CONTIGUOUS"]
        return C.c_void_p(v.ctypes.data)
    if hasattr(v, "data_ptr"):
        return C.c_void_p(v.data_ptr())
    return C.c_void_p(int(v))


class Lib:
    """One context of a library exporting the mdb200 call shapes under `prefix`."""

    def __init__(self, cdll, prefix, *create_args):
        self.cdll = cdll
        self.prefix = prefix
        create = getattr(cdll, prefix + "create")
        create.restype = C.c_void_p
        self._last_error = getattr(cdll, prefix + "last_error")
        self._last_error.restype = C.c_char_p
        self._last_error.argtypes = [C.c_void_p]
        self.ctx = create(*create_args)
        if not self.ctx:
            raise MdbError("%screate failed: %s" % (prefix, (self._last_error(None) or b"").decode()))
        self.dim = 0
        self.ncells = 0
        self.ndof = 0
        self.nchildren = 0

    # -- helpers --
    def _fn(self, name, restype=C.c_int):
        f = getattr(self.cdll, self.prefix + name)
        f.restype = restype
        return f

    def _check(self, rc, what):
        if rc < 0:
            raise MdbError("%s%s failed (%d): %s" % (self.prefix, what, rc, self._last_error(C.c_void_p(self.ctx)).decode()))
        return rc

    def close(self):
        if self.ctx:
            d = self._fn("destroy", None)
            d(C.c_void_p(self.ctx))
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- mesh --
    def mesh_structured(self, n, lower=None, upper=None):
        dim = len(n)
        lower = [0.0] * dim if lower is None else lower
        upper = [1.0] * dim if upper is None else upper
        nn, lo, up = _i64(n), np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
        self._check(self._fn("mesh_structured")(C.c_void_p(self.ctx), dim, _ptr(nn, C.c_int64), _ptr(lo, C.c_double),
                                                _ptr(up, C.c_double)), "mesh_structured")
        self.dim, self.n, self.ncells = dim, tuple(int(v) for v in n), int(np.prod(nn))
        self.nchildren = 0

    def mesh_unstructured(self, xy, conn):
        """2-D triangle mesh: xy (nv, 2) float64, conn (ne, 3) vertex indices (mdb_mesh_unstructured)."""
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        cn = np.ascontiguousarray(conn, dtype=np.int64)
        assert xy.ndim == 2 and xy.shape[1] == 2 and cn.ndim == 2 and cn.shape[1] == 3
        self._check(self._fn("mesh_unstructured")(C.c_void_p(self.ctx), 2, C.c_int64(xy.shape[0]), _ptr(xy, C.c_double),
                                                  C.c_int64(cn.shape[0]), ETYPE_TRIANGLE, _ptr(cn, C.c_int64)), "mesh_unstructured")
        self.dim, self.n, self.ncells = 2, (int(cn.shape[0]),), int(cn.shape[0])
        self.nchildren = 0

    def mesh_partition(self, global_n, cell_offset, global_lower, global_upper):
        gn, off = _i64(global_n), _i64(cell_offset)
        lo, up = np.asarray(global_lower, dtype=np.float64), np.asarray(global_upper, dtype=np.float64)
        self._check(self._fn("mesh_partition")(C.c_void_p(self.ctx), _ptr(gn, C.c_int64), _ptr(off, C.c_int64),
                                               _ptr(lo, C.c_double), _ptr(up, C.c_double)), "mesh_partition")

    def subdomains(self, bits):
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        assert b.size == self.ncells
        self._check(self._fn("subdomains")(C.c_void_p(self.ctx), _ptr(b, C.c_uint8)), "subdomains")

    def cell_groups(self, groups):
        """elementIndexToPhysicalGroup of the Gmsh reader (test/stokesdarcy2.cc:577-578): one int32 per cell."""
        g = np.ascontiguousarray(groups, dtype=np.int32)
        assert g.size == self.ncells
        self._check(self._fn("cell_groups")(C.c_void_p(self.ctx), _ptr(g, C.c_int32)), "cell_groups")

    def get_subdomains(self):
        b = np.zeros(self.ncells, dtype=np.uint8)
        self._check(self._fn("get_subdomains")(C.c_void_p(self.ctx), _ptr(b, C.c_uint8)), "get_subdomains")
        return b

    def subdomains_halfspace(self, axis, threshold, sd_below, sd_above):
        self._check(self._fn("subdomains_halfspace")(C.c_void_p(self.ctx), axis, C.c_double(threshold), sd_below, sd_above),
                    "subdomains_halfspace")

    # -- spaces / participants --
    def add_space(self, fe_kind, subdomain):
        r = self._check(self._fn("add_space")(C.c_void_p(self.ctx), fe_kind, subdomain), "add_space")
        self.nchildren += 1
        return r

    def add_subproblem(self, op_id, params, cond_kind, cond_mask, children):
        ch = np.ascontiguousarray(children, dtype=np.int32)
        return self._check(self._fn("add_subproblem")(C.c_void_p(self.ctx), op_id, C.byref(params), C.c_size_t(C.sizeof(params)),
                                                      cond_kind, C.c_uint32(cond_mask), _ptr(ch, C.c_int), len(ch)), "add_subproblem")

    def add_coupling(self, op_id, params, local_sp, remote_sp, jac_mode=JAC_FD_BITMATCH):
        return self._check(self._fn("add_coupling")(C.c_void_p(self.ctx), op_id, C.byref(params), C.c_size_t(C.sizeof(params)),
                                                    local_sp, remote_sp, jac_mode), "add_coupling")

    def add_coupling_space(self, fe_kind, left_cond, right_cond):
        """left_cond / right_cond = (cond_kind, mask) of the SubProblemSubProblemInterface predicate."""
        r = self._check(self._fn("add_coupling_space")(C.c_void_p(self.ctx), fe_kind, left_cond[0], C.c_uint32(left_cond[1]),
                                                       right_cond[0], C.c_uint32(right_cond[1])), "add_coupling_space")
        self.nchildren += 1
        return r

    def add_power_coupling_space(self, fe_kind, left_cond, right_cond, ncomp):
        r = self._check(self._fn("add_power_coupling_space")(C.c_void_p(self.ctx), fe_kind, left_cond[0], C.c_uint32(left_cond[1]),
                                                             right_cond[0], C.c_uint32(right_cond[1]), ncomp), "add_power_coupling_space")
        self.nchildren += ncomp
        return r

    def add_power_enriched_coupling(self, op_id, params, local_sp, remote_sp, first_child, ncomp, jac_mode=JAC_FD_BITMATCH):
        return self._check(self._fn("add_power_enriched_coupling")(C.c_void_p(self.ctx), op_id, C.byref(params), C.c_size_t(C.sizeof(params)),
                                                                   local_sp, remote_sp, first_child, ncomp, jac_mode), "add_power_enriched_coupling")

    def add_enriched_coupling(self, op_id, params, local_sp, remote_sp, coupling_child, jac_mode=JAC_FD_BITMATCH):
        return self._check(self._fn("add_enriched_coupling")(C.c_void_p(self.ctx), op_id, C.byref(params), C.c_size_t(C.sizeof(params)),
                                                             local_sp, remote_sp, coupling_child, jac_mode), "add_enriched_coupling")

    def finalize(self):
        nd = C.c_int64(0)
        self._check(self._fn("finalize")(C.c_void_p(self.ctx), C.byref(nd)), "finalize")
        self.ndof = nd.value
        return self.ndof

    def get_dofmap(self):
        ptr = np.zeros(self.ncells + 1, dtype=np.int64)
        self._check(self._fn("get_dofmap")(C.c_void_p(self.ctx), _ptr(ptr, C.c_int64), None), "get_dofmap")
        dofs = np.zeros(int(ptr[-1]), dtype=np.int64)
        self._check(self._fn("get_dofmap")(C.c_void_p(self.ctx), _ptr(ptr, C.c_int64), _ptr(dofs, C.c_int64)), "get_dofmap")
        return ptr, dofs

    def get_child_offsets(self):
        off = np.zeros(self.nchildren + 1, dtype=np.int64)
        self._check(self._fn("get_child_offsets")(C.c_void_p(self.ctx), _ptr(off, C.c_int64)), "get_child_offsets")
        return off

    # -- constraints / interpolation --
    def constraints_assemble(self, subproblems, bcs):
        sp = np.ascontiguousarray(subproblems, dtype=np.int32)
        arr = (BCType * len(bcs))(*bcs)
        nc = C.c_int64(0)
        self._check(self._fn("constraints_assemble")(C.c_void_p(self.ctx), _ptr(sp, C.c_int), arr, len(bcs), C.byref(nc)),
                    "constraints_assemble")
        return nc.value

    def set_constraints(self, dofs):
        d = _i64(dofs)
        self._check(self._fn("set_constraints")(C.c_void_p(self.ctx), C.c_int64(d.size), _ptr(d, C.c_int64)), "set_constraints")

    def get_constraints(self):
        f = np.zeros(self.ndof, dtype=np.uint8)
        self._check(self._fn("get_constraints")(C.c_void_p(self.ctx), _ptr(f, C.c_uint8)), "get_constraints")
        return f

    def interpolate(self, subproblems, fns, x, time=0.0):
        sp = np.ascontiguousarray(subproblems, dtype=np.int32)
        arr = (Function * len(fns))(*fns)
        self._check(self._fn("interpolate")(C.c_void_p(self.ctx), _ptr(sp, C.c_int), arr, len(fns), C.c_double(time), _vecptr(x)),
                    "interpolate")
        return x

    def copy_nonconstrained(self, xold, xnew):
        self._check(self._fn("copy_nonconstrained")(C.c_void_p(self.ctx), _vecptr(xold), _vecptr(xnew)), "copy_nonconstrained")

    # -- operators --
    def gridoperator_create(self, subproblems, couplings=()):
        sp = np.ascontiguousarray(subproblems, dtype=np.int32)
        cp = np.ascontiguousarray(couplings, dtype=np.int32)
        return self._check(self._fn("gridoperator_create")(C.c_void_p(self.ctx), _ptr(sp, C.c_int), len(sp), _ptr(cp, C.c_int), len(cp)),
                           "gridoperator_create")

    def matrix_create(self, gos):
        g = np.ascontiguousarray(gos, dtype=np.int32)
        nnz = C.c_int64(0)
        m = self._check(self._fn("matrix_create")(C.c_void_p(self.ctx), _ptr(g, C.c_int), len(g), C.byref(nnz)), "matrix_create")
        self.nnz = getattr(self, "nnz", {})
        self.nnz[m] = nnz.value
        return m

    def matrix_get_pattern(self, mat):
        rowptr = np.zeros(self.ndof + 1, dtype=np.int64)
        colidx = np.zeros(self.nnz[mat], dtype=np.int64)
        self._check(self._fn("matrix_get_pattern")(C.c_void_p(self.ctx), mat, _ptr(rowptr, C.c_int64), _ptr(colidx, C.c_int64)),
                    "matrix_get_pattern")
        return rowptr, colidx

    def matrix_get_values(self, mat, out=None):
        v = np.zeros(self.nnz[mat], dtype=np.float64) if out is None else out
        self._check(self._fn("matrix_get_values")(C.c_void_p(self.ctx), mat, _vecptr(v)), "matrix_get_values")
        return v

    def matrix_zero(self, mat):
        self._check(self._fn("matrix_zero")(C.c_void_p(self.ctx), mat), "matrix_zero")

    def set_time(self, t):
        self._check(self._fn("set_time")(C.c_void_p(self.ctx), C.c_double(t)), "set_time")

    def residual(self, go, x, r, weight=1.0):
        self._check(self._fn("residual")(C.c_void_p(self.ctx), go, C.c_double(weight), _vecptr(x), _vecptr(r)), "residual")
        return r

    def spmv(self, mat, x, y):
        self._check(self._fn("spmv")(C.c_void_p(self.ctx), mat, _vecptr(x), _vecptr(y)), "spmv")
        return y


class Mdb(Lib):
    """A context of libmdb200.so (the CUDA backend)."""

    def __init__(self, device=0):
        super().__init__(load_library(), "mdb_", C.c_int(device))

    def jacobian(self, go, mat, x, weight=1.0, overwrite=False):
        self._check(self._fn("jacobian")(C.c_void_p(self.ctx), go, mat, C.c_double(weight), _vecptr(x), int(bool(overwrite))), "jacobian")

    def solve(self, mat, rhs, z, method=SOLVER_CG, precond=PRECOND_JACOBI, reduction=1e-10, maxit=5000, fixed_iters=None):
        st = SolveStats()
        if fixed_iters is None:
            rc = self._fn("solve")(C.c_void_p(self.ctx), mat, method, precond, C.c_double(reduction), maxit,
                                   _vecptr(rhs), _vecptr(z), C.byref(st))
        else:
            rc = self._fn("solve_fixed")(C.c_void_p(self.ctx), mat, method, precond, int(fixed_iters),
                                         _vecptr(rhs), _vecptr(z), C.byref(st))
        if rc != -6:  # MDB_ENOTCONVERGED is reported through stats
            self._check(rc, "solve")
        return st

    def solve_block(self, mat, child, rhs, z, method=SOLVER_BICGSTAB, precond=PRECOND_SSOR, sweeps=3, reduction=1e-10, maxit=5000):
        """mdb_solve_block: the diagonal block of child space `child` only (ISTL_SEQ_Subblock_Backend)."""
        st = SolveStats()
        if precond == PRECOND_SSOR:
            self.set_ssor_sweeps(sweeps)
        rc = self._fn("solve_block")(C.c_void_p(self.ctx), mat, child, method, precond, C.c_double(reduction), maxit, _vecptr(rhs), _vecptr(z), C.byref(st))
        if rc != -6:
            self._check(rc, "solve_block")
        return st

    def onestep_apply(self, go0, go1, mat, subproblems, fns, time, dt, u_old, u_new, scheme=0, theta=1.0, method=SOLVER_BICGSTAB,
                      precond=PRECOND_JACOBI, reduction=1e-10, maxit=5000):
        """One time step on the device (mdb_onestep_apply); `time` is the time at the START of the step (OneStepMethod::apply)."""
        st = SolveStats()
        sp = np.ascontiguousarray(subproblems, dtype=np.int32)
        arr = (Function * len(fns))(*fns)
        self._check(self._fn("onestep_apply")(C.c_void_p(self.ctx), go0, go1, mat, _ptr(sp, C.c_int), arr, len(fns), C.c_double(time), C.c_double(dt),
                                              int(scheme), C.c_double(theta), method, precond, C.c_double(reduction), maxit, _vecptr(u_old), _vecptr(u_new), C.byref(st)),
                    "onestep_apply")
        return st

    def stationary_solve(self, go, mat, x, method=SOLVER_CG, precond=PRECOND_JACOBI, reduction=1e-10, maxit=5000, keep_matrix=False):
        st = SolveStats()
        rc = self._fn("stationary_solve")(C.c_void_p(self.ctx), go, mat, method, precond, C.c_double(reduction), maxit, _vecptr(x), int(keep_matrix),
                                          C.byref(st))
        self._check(rc, "stationary_solve")
        return st

    def set_halo_timeout(self, seconds):
        self._check(self._fn("set_halo_timeout")(C.c_void_p(self.ctx), C.c_double(seconds)), "set_halo_timeout")

    def mg_setup(self, mat):
        """mdb_mg_setup: builds the multigrid hierarchy of `mat` (collective on a partition); returns the number of levels."""
        n = C.c_int(0)
        self._check(self._fn("mg_setup")(C.c_void_p(self.ctx), mat, C.byref(n)), "mg_setup")
        return n.value

    def mg_level(self, mat, level):
        """The context of a coarse multigrid level as a borrowed binding (never destroyed from here): for halo_export / halo_import."""
        f = self._fn("mg_level", C.c_void_p)
        ptr = f(C.c_void_p(self.ctx), mat, level)
        if not ptr:
            raise MdbError("mdb_mg_level: no such level")
        b = object.__new__(Mdb)
        b.cdll, b.prefix, b.ctx, b._last_error = self.cdll, self.prefix, ptr, self._last_error
        b.dim, b.ncells, b.ndof, b.nchildren = self.dim, 0, 0, self.nchildren
        b.close = lambda: None                                  # borrowed: the fine context owns it
        return b

    def matrix_mg_levels(self, mat):
        return self._check(self._fn("matrix_mg_levels")(C.c_void_p(self.ctx), mat), "matrix_mg_levels")

    def set_ssor_sweeps(self, sweeps):
        self._check(self._fn("set_ssor_sweeps")(C.c_void_p(self.ctx), int(sweeps)), "set_ssor_sweeps")

    def matrix_device_ptrs(self, mat):
        rp, ci, va = C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)
        self._check(self._fn("matrix_device_ptrs")(C.c_void_p(self.ctx), mat, C.byref(rp), C.byref(ci), C.byref(va)), "matrix_device_ptrs")
        return rp.value, ci.value, va.value

    def phase_ms(self, phase):
        return self._fn("phase_ms", C.c_double)(C.c_void_p(self.ctx), phase.encode())

    def phase_count(self, phase):
        return self._fn("phase_count")(C.c_void_p(self.ctx), phase.encode())

    def profile_reset(self):
        self._fn("profile_reset")(C.c_void_p(self.ctx))

    def matrix_norm2(self, mat):
        out = C.c_double(0.0)
        self._check(self._fn("matrix_norm2")(C.c_void_p(self.ctx), mat, C.byref(out)), "matrix_norm2")
        return out.value

    def matrix_stats(self, mat):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._fn("matrix_stats")(C.c_void_p(self.ctx), mat, C.byref(a), C.byref(b)), "matrix_stats")
        return a.value, b.value

    def bytes_copied(self, reset=False):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._fn("bytes_copied")(C.c_void_p(self.ctx), C.byref(a), C.byref(b), int(reset)), "bytes_copied")
        return a.value, b.value

    def launch_count(self, reset=False):
        return self._fn("launch_count", C.c_int64)(C.c_void_p(self.ctx), int(reset))

    def stream(self):
        return self._fn("stream", C.c_void_p)(C.c_void_p(self.ctx))

    def synchronize(self):
        self._check(self._fn("synchronize")(C.c_void_p(self.ctx)), "synchronize")

    # multi-GPU plumbing
    def nccl_unique_id(self):
        buf = (C.c_char * 128)()
        self._check(self._fn("nccl_unique_id")(buf), "nccl_unique_id")
        return bytes(buf)

    def comm_init(self, nranks, rank, uid):
        self._check(self._fn("comm_init")(C.c_void_p(self.ctx), nranks, rank, C.c_char_p(uid)), "comm_init")

    def halo_export(self):
        buf = (C.c_char * 80)()
        nbytes = C.c_int64(0)
        self._check(self._fn("halo_export")(C.c_void_p(self.ctx), buf, C.byref(nbytes)), "halo_export")
        return bytes(buf), nbytes.value

    def scalars_export(self):
        """mdb_scalars_export: descriptor of this rank's scalar window (peer-memory all-reduce of the Krylov scalars)."""
        buf = (C.c_char * 80)()
        self._check(self._fn("scalars_export")(C.c_void_p(self.ctx), buf), "scalars_export")
        return bytes(buf)

    def scalars_exchanges(self):
        return int(self._fn("scalars_exchanges", C.c_int64)(C.c_void_p(self.ctx)))

    def scalars_import(self, blobs):
        buf = (C.c_char * (80 * max(len(blobs), 1)))()
        for i, b in enumerate(blobs):
            buf[80 * i:80 * (i + 1)] = b
        self._check(self._fn("scalars_import")(C.c_void_p(self.ctx), len(blobs), buf), "scalars_import")       # an empty list disconnects

    def halo_import(self, lower_rank, lower_blob, upper_rank, upper_blob):
        self._check(self._fn("halo_import")(C.c_void_p(self.ctx), lower_rank, C.c_char_p(lower_blob) if lower_blob else None,
                                            upper_rank, C.c_char_p(upper_blob) if upper_blob else None), "halo_import")

    def partition_lists(self, owned, neigh, send_ptr, send_idx, recv_ptr, recv_idx):
        """mdb_partition_lists: ownership flags and per-neighbour send / receive index lists of an unstructured partition."""
        o = np.ascontiguousarray(owned, dtype=np.uint8)
        nb = np.ascontiguousarray(neigh, dtype=np.int32)
        sp, rp = _i64(send_ptr), _i64(recv_ptr)
        si, ri = np.ascontiguousarray(send_idx, dtype=np.int32), np.ascontiguousarray(recv_idx, dtype=np.int32)
        assert o.size == self.ndof and sp.size == nb.size + 1 and rp.size == nb.size + 1
        self._check(self._fn("partition_lists")(C.c_void_p(self.ctx), _ptr(o, C.c_uint8), int(nb.size), _ptr(nb, C.c_int), _ptr(sp, C.c_int64),
                                                _ptr(si, C.c_int32), _ptr(rp, C.c_int64), _ptr(ri, C.c_int32)), "partition_lists")

    def halo_import_lists(self, blobs, caps, offsets, slots):
        buf = (C.c_char * (80 * max(len(blobs), 1)))()
        for i, b in enumerate(blobs):
            buf[80 * i:80 * (i + 1)] = bytes(b)
        cp, of, sl = _i64(caps), _i64(offsets), np.ascontiguousarray(slots, dtype=np.int32)
        self._check(self._fn("halo_import_lists")(C.c_void_p(self.ctx), buf, _ptr(cp, C.c_int64), _ptr(of, C.c_int64), _ptr(sl, C.c_int)), "halo_import_lists")

    def halo_push(self, x):
        self._check(self._fn("halo_push")(C.c_void_p(self.ctx), _vecptr(x)), "halo_push")

    def halo_wait(self, x):
        self._check(self._fn("halo_wait")(C.c_void_p(self.ctx), _vecptr(x)), "halo_wait")

    def owned_rows(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._fn("owned_rows")(C.c_void_p(self.ctx), C.byref(a), C.byref(b)), "owned_rows")
        return a.value, b.value

    def owned_counts(self):
        out = np.zeros(self.nchildren, dtype=np.int64)
        self._check(self._fn("owned_counts")(C.c_void_p(self.ctx), _ptr(out, C.c_int64)), "owned_counts")
        return out

    def get_ownership(self, child_global_base=None):
        flags = np.zeros(self.ndof, dtype=np.uint8)
        gidx = np.zeros(self.ndof, dtype=np.int64)
        base = None if child_global_base is None else _i64(child_global_base)
        self._check(self._fn("get_ownership")(C.c_void_p(self.ctx), _ptr(base, C.c_int64) if base is not None else None,
                                              _ptr(flags, C.c_uint8), _ptr(gidx, C.c_int64)), "get_ownership")
        return flags, gidx


_LIB = None


def library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmdb200.so")


def load_library():
    """Load libmdb200.so from the package directory.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is None:
        path = library_path()
        if not os.path.exists(path):
            raise MdbError("libmdb200.so not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C dune-multidomain_b200/csrc`; there is no CPU fallback" % path)
        _LIB = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _LIB
