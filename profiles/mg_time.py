"""profiles/mg_time.py — time to solution of the M2 system (3-D Q1/Q1 two-subdomain Poisson + CVCF coupling) and of one implicit
Euler step of M3 (2-D, mass + dt stiffness + proportional-flow coupling): CG / BiCGSTAB with Jacobi against the multigrid V-cycle
(MDB_PRECOND_MG).  Usage: python profiles/mg_time.py [3-D size] [2-D size]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
S3 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S2 = int(sys.argv[2]) if len(sys.argv) > 2 else 4096

dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (S3, S3, S3))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); z = torch.zeros_like(u)
dev.interpolate(P.sps, P.gfn, u); dev.jacobian(P.go, P.mat, u, overwrite=True); dev.residual(P.go, u, r)
out = {"M2": {"cells": S3 ** 3, "dofs": P.ndof}}
for name, method, precond in (("cg_jacobi", capi.SOLVER_CG, capi.PRECOND_JACOBI), ("cg_mg", capi.SOLVER_CG, capi.PRECOND_MG), ("cg_mg_again", capi.SOLVER_CG, capi.PRECOND_MG),
                              ("bicgstab_mg", capi.SOLVER_BICGSTAB, capi.PRECOND_MG)):
    dev.synchronize(); t0 = time.perf_counter()
    st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-10, maxit=20000)
    dev.synchronize(); wall = time.perf_counter() - t0
    y = torch.zeros_like(u); dev.spmv(P.mat, z, y); dev.synchronize()
    out["M2"][name] = {"iterations": st.iterations, "converged": bool(st.converged), "krylov_s": st.seconds, "wall_s": wall, "rel_residual": float((y - r).norm() / r.norm()),
                       "levels": dev.matrix_mg_levels(P.mat), "setup_ms": dev.phase_ms("mg_setup") * max(dev.phase_count("mg_setup"), 0)}
print(json.dumps(out), flush=True)
dev.close()

dev = pkg.Mdb(0)
P = problems.instationary(dev, (S2, S2))
u0 = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); u1 = torch.zeros_like(u0)
dev.interpolate(P.sps, P.gfn, u0, time=0.0)
out = {"M3": {"cells": S2 * S2, "dofs": P.ndof, "dt": 0.04 / 200}}
dt = 0.04 / 200
for name, method, precond in (("bicgstab_jacobi", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI), ("bicgstab_mg", capi.SOLVER_BICGSTAB, capi.PRECOND_MG), ("cg_mg", capi.SOLVER_CG, capi.PRECOND_MG)):
    steps = []
    ua = u0.clone()
    for k in range(1, 4):
        dev.synchronize(); t0 = time.perf_counter()
        st = dev.onestep_apply(P.go0, P.go1, P.mat, P.sps, P.gfn, (k - 1) * dt, dt, ua, u1, scheme=capi.ONESTEP_IMPLICIT_EULER, method=method, precond=precond, reduction=1e-10, maxit=20000)
        dev.synchronize(); steps.append({"iterations": st.iterations, "krylov_s": st.seconds, "wall_s": time.perf_counter() - t0})
        ua, u1 = u1, ua
    out["M3"][name] = {"steps": steps, "checksum": float(ua.sum())}
print(json.dumps(out), flush=True)
dev.close()
