"""profiles/mortar_time.py — BASELINE configs[3] at size: testmortarpoisson (Q1 | Q1 + Q1 mortar space on the interface, EnrichedCoupling with
MortarPoissonCoupling and the reference's FD Jacobians) on a 3-D structured grid.  Usage: python profiles/mortar_time.py [cells per axis ...]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
for n in [int(a) for a in sys.argv[1:]] or [128, 256]:
    for ncomp in (1, 2):
        dev = pkg.Mdb(0)
        t0 = time.perf_counter()
        P = problems.testmortarpoisson(dev, (n, n, n), ncomp=ncomp)
        dev.synchronize(); t_setup = time.perf_counter() - t0
        u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); y = torch.zeros_like(u)
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
        nnz = dev.nnz[P.mat]
        print(json.dumps({"cells": n ** 3, "driver": "testmortarpoisson" if ncomp == 1 else "testpowermortarpoisson", "dofs": int(P.ndof), "nnz": int(nnz),
                          "mortar_dofs": int(P.ndof - sum(dev.get_child_offsets()[2:3]) if False else 0) or None, "setup_s": t_setup,
                          "jacobian_ms": ms_j, "jacobian_GDOFs": P.ndof / ms_j / 1e6, "jacobian_GBs_8nnz": 8 * nnz / ms_j / 1e6,
                          "residual_ms": ms_r, "spmv_ms": ms_s, "spmv_GBs_12nnz": 12 * nnz / ms_s / 1e6}), flush=True)
        dev.close()
