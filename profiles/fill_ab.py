"""A/B of the simple-row fill kernels on a single-domain 256^3 Poisson Jacobian (no coupling: the step is elem + fill + boundary rows).
MDB_FILL_NO_TMA=1 selects the per-thread store kernel, MDB_FILL_TMA_BLOCKS_PER_SM the grid of the bulk-store kernel."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = pkg.Mdb(0)
P = problems.single_domain_poisson(dev, (N, N, N))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda")
for _ in range(3): dev.jacobian(P.go, P.mat, u, overwrite=True)
dev.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): dev.jacobian(P.go, P.mat, u, overwrite=True)
    dev.synchronize()
    best = min(best, (time.perf_counter() - t0) / 10 * 1e3)
nnz = dev.nnz[P.mat]
print("no_tma=%s blocks/SM=%s: jacobian %.4f ms  (%.0f GB/s of matrix values)  fill phase %.4f ms" % (os.environ.get("MDB_FILL_NO_TMA", "0"),
      os.environ.get("MDB_FILL_TMA_BLOCKS_PER_SM", "2"), best, 8.0 * nnz / best * 1e-6, dev.phase_ms("jacobian_fill")))
