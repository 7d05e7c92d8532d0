"""Times the QkDG / ConvectionDiffusionDG path (BASELINE configs[0] as shipped: DG-Q2, SIPG, CVCF coupling) on one B200:
python profiles/dg_time.py [refine ...]   (the reference's ini: mesh.refine = 9).  Device vectors, CUDA-event phases of the context."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
refines = [int(a) for a in sys.argv[1:]] or [7, 9]
for refine in refines:
    n = (1 << refine,) * 2
    dev = pkg.Mdb(0)
    t0 = time.perf_counter()
    P = problems.dirichlet_neumann_dg(dev, n)
    dev.synchronize(); t_setup = time.perf_counter() - t0
    u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); z = torch.zeros_like(u)
    nnz = dev.nnz[P.mat]
    def timed(fn, reps=int(os.environ.get("DG_TIME_REPS", "20"))):
        for _ in range(3): fn()
        dev.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(reps): fn()
        dev.synchronize()
        return (time.perf_counter() - t) / reps * 1e3
    t_jac = timed(lambda: dev.jacobian(P.go, P.mat, u, overwrite=True))
    t_res = timed(lambda: (r.zero_(), dev.residual(P.go, u, r)))
    y = torch.zeros_like(u)
    t_spmv = timed(lambda: dev.spmv(P.mat, u, y))
    print("refine=%d ndof=%d nnz=%d setup(dofmap+pattern) %.1f ms | jacobian %.3f ms = %.2f GDOF/s, %.0f GB/s of values | residual %.3f ms | spmv %.3f ms = %.0f GB/s"
          % (refine, P.ndof, nnz, t_setup * 1e3, t_jac, P.ndof / t_jac / 1e6, 8 * nnz / t_jac / 1e6, t_res, t_spmv, (12 * nnz + 20 * P.ndof) / t_spmv / 1e6), flush=True)
    r.zero_(); dev.residual(P.go, u, r)
    if os.environ.get("DG_TIME_NO_SOLVE"):
        dev.close(); continue
    for name, method, precond, sweeps, maxit in (("BiCGSTAB+Jacobi", capi.SOLVER_BICGSTAB, capi.PRECOND_JACOBI, 1, 20000),
                                                 ("BiCGSTAB+SSOR(3)", capi.SOLVER_BICGSTAB, capi.PRECOND_SSOR, 3, 12 if refine > 7 else 5000),
                                                 ("BiCGSTAB+MG", capi.SOLVER_BICGSTAB, capi.PRECOND_MG, 1, 500), ("BiCGSTAB+MG again", capi.SOLVER_BICGSTAB, capi.PRECOND_MG, 1, 500)):
        if os.environ.get("DG_TIME_ONLY") and os.environ["DG_TIME_ONLY"] not in name: continue
        dev.set_ssor_sweeps(sweeps)
        st = dev.solve(P.mat, r, z, method=method, precond=precond, reduction=1e-8, maxit=maxit)
        dev.synchronize()
        print("   %-18s iters=%5d converged=%d  loop %.1f ms  (%.3f ms/iter)" % (name, st.iterations, st.converged, st.seconds * 1e3,
              st.seconds * 1e3 / max(st.iterations, 1)), flush=True)
    dev.close()
