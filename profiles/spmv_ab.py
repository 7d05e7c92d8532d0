import importlib, os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200")
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (256, 256, 256))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); y = torch.zeros_like(u)
dev.interpolate(P.sps, P.gfn, u)
dev.jacobian(P.go, P.mat, u, overwrite=True)
for _ in range(3): dev.spmv(P.mat, u, y)
dev.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(20): dev.spmv(P.mat, u, y)
    dev.synchronize()
    best = min(best, (time.perf_counter() - t0) / 20 * 1e3)
print("dbg", os.environ.get("MDB_SPMV_DBG", "0"), "variant", os.environ.get("MDB_SPMV_VARIANT", "4"), "spmv ms %.4f (wall clock over 20 back-to-back launches)" % best)
