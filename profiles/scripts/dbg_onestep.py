"""Diagnostic for tests/test_gpu_um.py::test_one_step_theta_on_a_triangle_mesh: per-step difference device vs oracle stage."""
import sys, importlib, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scipy.sparse as sps, scipy.sparse.linalg as spla
import um_problems as ump
pkg = importlib.import_module("dune-multidomain_b200"); capi = pkg.capi
for theta in (1.0, 0.5):
    xy, conn = ump.square_mesh(9, 8, seed=21); sd = ump.halves(xy, conn)
    dev, ora = pkg.Mdb(0), ump.oracle()
    Hd = ump.setup(dev, xy, conn, sd, with_mass=True, cvcf=True, k=2.0)
    Ho = ump.setup(ora, xy, conn, sd, with_mass=True, cvcf=True, k=2.0)
    g = capi.Function(capi.FN_INSTAT_G, 1.5, 3.0)
    n, dt, t = Ho["ndof"], 0.05, 0.1
    ud = np.random.default_rng(3).uniform(-1.0, 1.0, n); uo = ud.copy()
    free = ora.get_constraints() == 0
    rp, ci = ora.matrix_get_pattern(Ho["mat"])
    for step in range(2):
        new = np.zeros(n)
        st = dev.onestep_apply(Hd["go"], Hd["go1"], Hd["mat"], list(Hd["sp"]), [g, g], t, dt, ud, new, scheme=capi.ONESTEP_THETA, theta=theta,
                               method=capi.SOLVER_BICGSTAB, precond=capi.PRECOND_SSOR, reduction=1e-13, maxit=2000)
        ud = new
        x = ora.interpolate(list(Ho["sp"]), [g, g], np.zeros(n), time=t + dt); x[free] = uo[free]
        ora.jacobian(Ho["go1"], Ho["mat"], x, weight=1.0, overwrite=True); ora.jacobian(Ho["go"], Ho["mat"], x, weight=theta * dt)
        R = np.zeros(n); ora.set_time(t); ora.residual(Ho["go1"], uo, R, weight=-1.0)
        if theta != 1.0: ora.residual(Ho["go"], uo, R, weight=(1.0 - theta) * dt)
        ora.set_time(t + dt); ora.residual(Ho["go1"], x, R, weight=1.0); ora.residual(Ho["go"], x, R, weight=theta * dt)
        A = sps.csr_matrix((ora.matrix_get_values(Ho["mat"]), ci, rp), shape=(n, n))
        uo = x - spla.spsolve(A.tocsc(), R)
        e = np.abs(ud - uo); i = int(e.argmax())
        print("theta", theta, "step", step, "iters", st.iterations, st.converged, "diff", e.max() / np.abs(uo).max(), "at", i, "free" if free[i] else "constrained", ud[i], uo[i], flush=True)
        t += dt
    dev.close()
