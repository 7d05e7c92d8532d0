"""Short driver for ncu captures of the unstructured path: 1024^2 (or argv[1]^2) jittered squares, P1 | P1, Poisson + ProportionalFlowCoupling."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rng = np.random.default_rng(0)
i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="xy")
xy = np.stack([i.ravel() / n, j.ravel() / n], axis=1)
inner = (i.ravel() > 0) & (i.ravel() < n) & (j.ravel() > 0) & (j.ravel() < n) & (i.ravel() != n // 2)
xy[inner] += rng.uniform(-0.15, 0.15, (int(inner.sum()), 2)) / n
ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
a = (cj * (n + 1) + ci).ravel(); b = a + 1; c = a + n + 1; d = c + 1
even = ((ci + cj) % 2 == 0).ravel()
conn = np.empty((2 * n * n, 3), dtype=np.int64)
conn[0::2] = np.where(even[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1)); conn[1::2] = np.where(even[:, None], np.stack([a, d, c], 1), np.stack([b, d, c], 1))
sd = np.where(xy[conn].mean(axis=1)[:, 0] > 0.5, 2, 1).astype(np.uint8)
dev = pkg.Mdb(0)
dev.mesh_unstructured(xy, conn); dev.subdomains(sd)
c0, c1 = dev.add_space(capi.FE_P1, 0), dev.add_space(capi.FE_P1, 1)
p = capi.PoissonParams(); p.f = capi.Function(capi.FN_ZERO); p.bc = capi.BCType(capi.BC_TESTPOISSON); p.j = capi.Function(capi.FN_TESTPOISSON_J); p.quad_order = 4
s0 = dev.add_subproblem(capi.OP_POISSON, p, capi.COND_EQUALITY, 1, [c0]); s1 = dev.add_subproblem(capi.OP_POISSON, p, capi.COND_EQUALITY, 2, [c1])
pf = capi.PropFlowParams(); pf.intensity = 2.0
cp = dev.add_coupling(capi.OP_PROPFLOW_COUPLING, pf, s0, s1)
ndof = dev.finalize(); dev.constraints_assemble([s0, s1], [p.bc, p.bc]); go = dev.gridoperator_create([s0, s1], [cp]); mat = dev.matrix_create([go])
u = torch.zeros(ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); y = torch.zeros_like(u)
dev.interpolate([s0, s1], [capi.gauss(), capi.gauss()], u)
for _ in range(reps):
    dev.jacobian(go, mat, u, overwrite=True)
    dev.residual(go, u, r)
    dev.spmv(mat, u, y)
dev.synchronize()
