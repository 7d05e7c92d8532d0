"""Phase times of the Jacobian step of BASELINE configs[3] (testmortarpoisson, 256^3): python profiles/mortar_phases.py [n] [ncomp]"""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ncomp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = pkg.Mdb(0)
P = problems.testmortarpoisson(dev, (n, n, n), ncomp=ncomp)
u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda")
dev.interpolate(P.sps, P.gfn, u)
for _ in range(3):
    dev.jacobian(P.go, P.mat, u, overwrite=True)
dev.synchronize(); dev.profile_reset()
s = torch.cuda.ExternalStream(dev.stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10):
    dev.jacobian(P.go, P.mat, u, overwrite=True)
e1.record(s); e1.synchronize()
out = {"ms": e0.elapsed_time(e1) / 10, "serial": os.environ.get("MDB_JAC_SERIAL")}
out.update({k: dev.phase_ms("jacobian_" + k) for k in ("elem", "fill", "rows", "coupling", "dirichlet")})
st = dev.matrix_stats(P.mat) if hasattr(dev, "matrix_stats") else None
out["stats"] = st
print(json.dumps(out))
dev.close()
