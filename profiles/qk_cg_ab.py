import importlib, os, sys, json
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, problems
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
n = 2048
for fe, name in (((capi.FE_Q2, capi.FE_Q2), "Q2|Q2"), ((capi.FE_Q1, capi.FE_Q2), "Q1|Q2")):
    dev = pkg.Mdb(0)
    P = problems.testpoisson(dev, (n, n), fe=fe)
    u = torch.zeros(P.ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); z = torch.zeros_like(u)
    dev.interpolate(P.sps, P.gfn, u)
    dev.jacobian(P.go, P.mat, u, overwrite=True)
    dev.residual(P.go, u, r)
    out = {}
    for it in (50, 50, 100, 200):
        st = dev.solve(P.mat, r, z, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, fixed_iters=it)
        out["cg%d" % it] = round(st.seconds * 1e3, 3)
    st = dev.solve(P.mat, r, z, method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_JACOBI, fixed_iters=100)
    out["bicgstab100"] = round(st.seconds * 1e3, 3)
    print(name, os.environ.get("MDB_SPMV_NO_RUNS"), json.dumps(out), flush=True)
    dev.close()
