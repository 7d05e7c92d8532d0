"""Short driver for ncu captures: sets up the M2 problem and runs every hot operation a couple of times."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import problems  # noqa: E402

pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (S, S, S))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda")
r = torch.zeros_like(u)
y = torch.zeros_like(u)
z = torch.zeros_like(u)
torch.cuda.synchronize()
dev.interpolate(P.sps, P.gfn, u)
for _ in range(reps):
    dev.jacobian(P.go, P.mat, u, overwrite=True)
    dev.residual(P.go, u, r)
    dev.spmv(P.mat, u, y)
dev.solve(P.mat, r, z, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, fixed_iters=3)
dev.synchronize()
print("ok", P.ndof, dev.nnz[P.mat])
