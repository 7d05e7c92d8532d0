import importlib, sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import orc, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
kw = dict(n=(36, 20), fe=(capi.FE_Q1, capi.FE_Q2), split_axis=0)
if len(sys.argv) > 1: kw = dict(n=(37, 18), fe=(capi.FE_Q1, capi.FE_Q2), split_axis=0, quad_order=6)
dev, ora = pkg.Mdb(0), orc.Oracle()
Pd, Po = problems.testpoisson(dev, **kw), problems.testpoisson(ora, **kw)
u0 = np.zeros(Po.ndof); ora.interpolate(Po.sps, Po.gfn, u0)
dev.jacobian(Pd.go, Pd.mat, u0, overwrite=True)
rd = np.zeros(Pd.ndof); dev.residual(Pd.go, u0, rd)
for red in (1e-8, 1e-10, 1e-11, 1e-12, 1e-13):
    for m in (capi.SOLVER_CG, capi.SOLVER_BICGSTAB):
        zd = np.zeros(Pd.ndof)
        sd = dev.solve(Pd.mat, rd, zd, method=m, precond=capi.PRECOND_JACOBI, reduction=red, maxit=20000)
        print(red, m, sd.converged, sd.iterations, sd.defect / sd.defect0, np.isfinite(zd).all())
