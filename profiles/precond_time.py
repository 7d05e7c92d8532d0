"""Times one application of the sequential preconditioners (sync-free sweeps) and the Krylov loops that use them on the M2
problem: python profiles/precond_time.py [N]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (N, N, N))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); z = torch.zeros_like(u)
dev.interpolate(P.sps, P.gfn, u)
dev.jacobian(P.go, P.mat, u, overwrite=True)
dev.residual(P.go, u, r)
for name, method, precond, sweeps in (("BiCGSTAB+SSOR(3)", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR, 3), ("CG+SSOR(1)", capi.SOLVER_CG, capi.PRECOND_SSOR, 1),
                                      ("BiCGSTAB+ILU0", capi.SOLVER_BICGSTAB, capi.PRECOND_ILU0, 1), ("CG+ILU0", capi.SOLVER_CG, capi.PRECOND_ILU0, 1),
                                      ("CG+Jacobi", capi.SOLVER_CG, capi.PRECOND_JACOBI, 1), ("BiCGSTAB+Jacobi", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI, 1)):
    dev.set_ssor_sweeps(sweeps)
    t0 = time.perf_counter()
    st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-8, maxit=5000)
    dev.synchronize()
    wall = time.perf_counter() - t0
    print("N=%d ndof=%d %-18s iters=%4d converged=%d  loop %.1f ms  (%.2f ms/iter)  wall %.1f ms" % (N, P.ndof, name, st.iterations, st.converged,
          st.seconds * 1e3, st.seconds * 1e3 / max(st.iterations, 1), wall * 1e3))
