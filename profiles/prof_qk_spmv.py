"""Driver for ncu: the Qk (k = 2) matrix-vector product on 2048^2 cells, Q2 | Q2 (profiles/qk_time.py's case): a few products only.
ncu --set full -k regex:k_spmv_runs -s 2 -c 2 python profiles/prof_qk_spmv.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (n, n), fe=(capi.FE_Q2, capi.FE_Q2))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); y = torch.zeros_like(u)
dev.interpolate(P.sps, P.gfn, u)
dev.jacobian(P.go, P.mat, u, overwrite=True)
for _ in range(5):
    dev.spmv(P.mat, u, y)
dev.synchronize()
print("ok", P.ndof, dev.nnz[P.mat])
dev.close()
