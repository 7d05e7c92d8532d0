"""Times mdb_residual on the M2 problem (3-D Q1/Q1 two-subdomain Poisson + CVCF coupling): python profiles/res_time.py [cells per axis]
Switches: MDB_RESIDUAL_NO_MARCH=1 (round 1's tile kernel in 3-D), MDB_RESIDUAL_ZT (planes per marching tile), MDB_RESIDUAL_SERIAL=1
(one stream: the phase time of the volume part is then the kernels' own)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (S, S, S))
u = torch.rand(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u)
for _ in range(5): dev.residual(P.go, u, r)
dev.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    torch.cuda.synchronize(); e0.record()
    import time; th = time.perf_counter()
    for _ in range(20): dev.residual(P.go, u, r)
    host = (time.perf_counter() - th) / 20 * 1e3
    dev.synchronize(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print("S=%d ndof=%d march=%s zt=%s serial=%s | mdb_residual %.4f ms = %.1f GDOF/s, 16 B/row -> %.0f GB/s | phase residual_volume %.4f ms | host time per call (no sync) %.4f ms"
      % (S, P.ndof, "0" if os.environ.get("MDB_RESIDUAL_NO_MARCH") else "1", os.environ.get("MDB_RESIDUAL_ZT", "16"), os.environ.get("MDB_RESIDUAL_SERIAL", "0"),
         best, P.ndof / best / 1e6, 16 * P.ndof / best / 1e6, dev.phase_ms("residual_volume"), host))
