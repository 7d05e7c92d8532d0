"""Times the unstructured-mesh path (csrc/unstructured.cu: triangles, P1 | P1, Poisson + ProportionalFlowCoupling) on one B200:
python profiles/um_time.py [squares_per_axis ...] — every square is cut into two triangles; cells in generation order ("sorted",
what a mesh generator with a space-filling order delivers) and in random order ("shuffled", the worst case for the gathers).
Device vectors; wall clock around synchronised repetitions."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi


def mesh(n, seed=0):
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="xy")
    xy = np.stack([i.ravel() / n, j.ravel() / n], axis=1)
    inner = (i.ravel() > 0) & (i.ravel() < n) & (j.ravel() > 0) & (j.ravel() < n) & (i.ravel() != n // 2)
    xy[inner] += rng.uniform(-0.15, 0.15, (int(inner.sum()), 2)) / n
    ci, cj = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    a = (cj * (n + 1) + ci).ravel(); b = a + 1; c = a + n + 1; d = c + 1
    even = ((ci + cj) % 2 == 0).ravel()
    t0 = np.where(even[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t1 = np.where(even[:, None], np.stack([a, d, c], 1), np.stack([b, d, c], 1))
    conn = np.empty((2 * n * n, 3), dtype=np.int64); conn[0::2] = t0; conn[1::2] = t1
    return np.ascontiguousarray(xy), conn


for n in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    xy, conn0 = mesh(n)
    for order in ("sorted", "shuffled"):
        conn = conn0 if order == "sorted" else np.ascontiguousarray(conn0[np.random.default_rng(1).permutation(conn0.shape[0])])
        sd = np.where(xy[conn].mean(axis=1)[:, 0] > 0.5, 2, 1).astype(np.uint8)
        dev = pkg.Mdb(0)
        t0 = time.perf_counter()
        dev.mesh_unstructured(xy, conn); dev.subdomains(sd)
        t_ingest = time.perf_counter() - t0
        c0, c1 = dev.add_space(capi.FE_P1, 0), dev.add_space(capi.FE_P1, 1)
        p = capi.PoissonParams(); p.f = capi.Function(capi.FN_ZERO); p.bc = capi.BCType(capi.BC_TESTPOISSON); p.j = capi.Function(capi.FN_TESTPOISSON_J); p.quad_order = 4
        s0 = dev.add_subproblem(capi.OP_POISSON, p, capi.COND_EQUALITY, 1, [c0]); s1 = dev.add_subproblem(capi.OP_POISSON, p, capi.COND_EQUALITY, 2, [c1])
        pf = capi.PropFlowParams(); pf.intensity = 2.0
        cp = dev.add_coupling(capi.OP_PROPFLOW_COUPLING, pf, s0, s1)
        t0 = time.perf_counter()
        ndof = dev.finalize(); dev.constraints_assemble([s0, s1], [p.bc, p.bc]); go = dev.gridoperator_create([s0, s1], [cp]); mat = dev.matrix_create([go])
        dev.synchronize(); t_setup = time.perf_counter() - t0
        nnz = dev.nnz[mat]
        u = torch.zeros(ndof, dtype=torch.float64, device="cuda"); r = torch.zeros_like(u); y = torch.zeros_like(u); z = torch.zeros_like(u)
        dev.interpolate([s0, s1], [capi.gauss(), capi.gauss()], u)

        def timed(fn, reps=20):
            for _ in range(3): fn()
            dev.synchronize(); t = time.perf_counter()
            for _ in range(reps): fn()
            dev.synchronize()
            return (time.perf_counter() - t) / reps * 1e3
        t_jac = timed(lambda: dev.jacobian(go, mat, u, overwrite=True))
        t_res = timed(lambda: (r.zero_(), dev.residual(go, u, r)))
        t_spmv = timed(lambda: dev.spmv(mat, u, y))
        ne = conn.shape[0]
        b_jac = 8 * nnz + 4 * 3 * ne + 4 * 3 * ne + 16 * xy.shape[0]      # values + DOF map (conn) + geometry (conn again, coordinates)
        print("n=%d %-8s cells=%d ndof=%d nnz=%d | ingest(host adjacency+upload) %.0f ms, dofmap+constraints+pattern %.1f ms | jacobian %.3f ms = %.2f GDOF/s, %.0f GB/s algorithmic"
              " | residual %.3f ms = %.2f GDOF/s | spmv %.3f ms = %.0f GB/s" % (n, order, ne, ndof, nnz, t_ingest * 1e3, t_setup * 1e3, t_jac, ndof / t_jac / 1e6,
              b_jac / t_jac / 1e6, t_res, ndof / t_res / 1e6, t_spmv, (12 * nnz + 20 * ndof) / t_spmv / 1e6), flush=True)
        for ph in ("jacobian_rows", "jacobian_coupling", "residual_volume", "residual_boundary", "residual_coupling"):
            print("      %-18s %.3f ms" % (ph, dev.phase_ms(ph)))
        r.zero_(); dev.residual(go, u, r)
        st = dev.solve(mat, r, z, method=capi.SOLVER_CG, precond=capi.PRECOND_JACOBI, reduction=1e-8, maxit=20000)
        print("      CG+Jacobi iters=%d converged=%d %.1f ms (%.3f ms/iter)" % (st.iterations, st.converged, st.seconds * 1e3, st.seconds * 1e3 / max(1, st.iterations)), flush=True)
        dev.close()
