"""Driver for ncu: the Jacobian of BASELINE configs[3] (testmortarpoisson, 256^3) a few times.
ncu --set full -k regex:k_mortar_jacobian -s 1 -c 1 python profiles/prof_mortar.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = pkg.Mdb(0)
P = problems.testmortarpoisson(dev, (n, n, n))
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda")
dev.interpolate(P.sps, P.gfn, u)
for _ in range(3):
    dev.jacobian(P.go, P.mat, u, overwrite=True)
dev.synchronize()
print("ok", P.ndof)
dev.close()
