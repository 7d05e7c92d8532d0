"""profiles/qk_time.py — the shipped testpoisson space combination (Q1 left | Q2 right, test/testpoisson.cc:168-198) at size:
Jacobian / residual / SpMV / CG times of the generic Qk kernels (csrc/generic.cu).  Usage: python profiles/qk_time.py [cells per axis ...]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
for n in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    for fe, name in (((capi.FE_Q1, capi.FE_Q2), "Q1|Q2"), ((capi.FE_Q2, capi.FE_Q2), "Q2|Q2"), ((capi.FE_Q1, capi.FE_Q1), "Q1|Q1")):
        dev = pkg.Mdb(0)
        t0 = time.perf_counter()
        P = problems.testpoisson(dev, (n, n), fe=fe)
        dev.synchronize(); t_setup = time.perf_counter() - t0
        u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); y = torch.zeros_like(u); z = torch.zeros_like(u)
        dev.interpolate(P.sps, P.gfn, u)
        s = torch.cuda.ExternalStream(dev.stream())
        def timed(fn, reps=10):
            for _ in range(3): fn()
            dev.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(reps): fn()
            e1.record(s); e1.synchronize()
            return e0.elapsed_time(e1) / reps
        ms_j = timed(lambda: dev.jacobian(P.go, P.mat, u, overwrite=True))
        ms_r = timed(lambda: dev.residual(P.go, u, r))
        ms_s = timed(lambda: dev.spmv(P.mat, u, y))
        r.zero_(); dev.residual(P.go, u, r)
        st = dev.solve(P.mat, r, z, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, fixed_iters=50)
        nnz = dev.nnz[P.mat]
        print(json.dumps({"cells": n * n, "spaces": name, "dofs": P.ndof, "nnz": nnz, "setup_s": t_setup, "jacobian_ms": ms_j, "jacobian_GDOFs": P.ndof / ms_j / 1e6,
                          "jacobian_GBs_8nnz": 8.0 * nnz / ms_j / 1e6, "residual_ms": ms_r, "spmv_ms": ms_s, "spmv_GBs_12nnz": 12.0 * nnz / ms_s / 1e6, "cg_ms_per_iter": st.seconds * 1e3 / 50}), flush=True)
        dev.close()
