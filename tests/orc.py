"""ctypes wrapper of the CPU oracle (oracle/_build/liboracle.so) — test infrastructure only."""
import ctypes as C
import importlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = importlib.import_module("dune-multidomain_b200").capi

_LIBS = {}


def load(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _LIBS:
        _LIBS[name] = C.CDLL(os.path.join(ROOT, "oracle", "_build", name))
    return _LIBS[name]


class Oracle(capi.Lib):
    def __init__(self, fast=False):
        super().__init__(load(fast), "orc_")

    def jacobian(self, go, mat, x, weight=1.0, overwrite=False, nthreads=0):
        if overwrite:
            self.matrix_zero(mat)
        if nthreads:
            self._check(self._fn("jacobian_parallel")(C.c_void_p(self.ctx), go, mat, C.c_double(weight), capi._vecptr(x), nthreads),
                        "jacobian_parallel")
        else:
            self._check(self._fn("jacobian")(C.c_void_p(self.ctx), go, mat, C.c_double(weight), capi._vecptr(x)), "jacobian")

    def solve(self, mat, rhs, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR_ORACLE, sweeps=3, reduction=1e-10, maxit=5000):
        st = capi.SolveStats()
        if precond == capi.PRECOND_SSOR:
            precond = capi.PRECOND_SSOR_ORACLE
        self._check(self._fn("solve")(C.c_void_p(self.ctx), mat, method, precond, sweeps, C.c_double(reduction), maxit,
                                      capi._vecptr(rhs), capi._vecptr(z), C.byref(st)), "solve")
        return st

    def solve_block(self, mat, child, rhs, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR_ORACLE, sweeps=3, reduction=1e-10, maxit=5000):
        st = capi.SolveStats()
        if precond == capi.PRECOND_SSOR:
            precond = capi.PRECOND_SSOR_ORACLE
        self._check(self._fn("solve_block")(C.c_void_p(self.ctx), mat, child, method, precond, sweeps, C.c_double(reduction), maxit,
                                            capi._vecptr(rhs), capi._vecptr(z), C.byref(st)), "solve_block")
        return st

    def local_jacobian_volume(self, sp, cell, nloc):
        a = np.zeros((nloc, nloc))
        self._check(self._fn("local_jacobian_volume")(C.c_void_p(self.ctx), sp, C.c_int64(cell), capi._vecptr(a)), "local_jacobian_volume")
        return a

    def csr(self, mat):
        import scipy.sparse as sp
        rowptr, colidx = self.matrix_get_pattern(mat)
        return sp.csr_matrix((self.matrix_get_values(mat), colidx, rowptr), shape=(self.ndof, self.ndof))
