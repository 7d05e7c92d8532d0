"""Short driver for ncu captures of mdb_residual on the M2 problem: python profiles/prof_res.py [cells per axis] [calls]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = pkg.Mdb(0)
P = problems.testpoisson(dev, (S, S, S))
u = torch.rand(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u)
for _ in range(reps): dev.residual(P.go, u, r)
dev.synchronize()
print("ok", P.ndof)
