"""profiles/sd_time.py — Stokes-Darcy path (csrc/ufe.cu, csrc/direct.cu) at size: phases of setup, assembly and the direct solve on
generated filtration meshes (tools/make_filtration_msh.py layout), CUDA-event times from mdb_phase_ms.  Usage: python profiles/sd_time.py [n ...]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi
import make_filtration_msh as mk  # noqa: E402
import sd_problems as sdp  # noqa: E402


def run(n, solve=True):
    xy, conn, groups = mk.build(nx=n, ny=n)
    dev = pkg.Mdb(0)
    t0 = time.perf_counter()
    H = sdp.setup(dev, xy, conn, groups.astype(np.int32))
    dev.synchronize()
    t_setup = time.perf_counter() - t0
    x = np.zeros(H["ndof"])
    out = dict(n=n, cells=int(conn.shape[0]), dofs=int(H["ndof"]), nnz=int(dev.nnz[H["mat"]]), setup_wall_s=t_setup,
               dofmap_ms=dev.phase_ms("dofmap"), pattern_ms=dev.phase_ms("pattern"))
    import torch
    xd = torch.zeros(H["ndof"], dtype=torch.float64, device="cuda"); rd = torch.zeros_like(xd)
    for _ in range(3):
        dev.jacobian(H["go"], H["mat"], xd, overwrite=True); dev.residual(H["go"], xd, rd)
    dev.profile_reset()
    for _ in range(10):
        dev.jacobian(H["go"], H["mat"], xd, overwrite=True); dev.residual(H["go"], xd, rd)
    dev.synchronize()
    for ph in ("jacobian_rows", "jacobian_coupling", "residual_volume", "residual_coupling"):
        out[ph + "_ms"] = dev.phase_ms(ph)
    out["jacobian_ms"] = out["jacobian_rows_ms"] + out["jacobian_coupling_ms"]
    out["jacobian_dofs_per_s"] = out["dofs"] / (out["jacobian_ms"] * 1e-3)
    out["jacobian_GBs"] = 8.0 * out["nnz"] / (out["jacobian_ms"] * 1e-3) / 1e9
    if solve:
        u = np.zeros(H["ndof"])
        dev.profile_reset()
        t0 = time.perf_counter()
        st = dev.stationary_solve(H["go"], H["mat"], u, method=capi.SOLVER_DIRECT, precond=capi.PRECOND_NONE)
        out["solve_wall_s"] = time.perf_counter() - t0
        out["direct_order_ms"] = dev.phase_ms("direct_order"); out["direct_factor_ms"] = dev.phase_ms("direct_factor"); out["direct_solve_ms"] = dev.phase_ms("direct_solve")
        out["defect_reduction"] = st.defect / st.defect0 if st.defect0 > 0 else 0.0
        out["converged"] = int(st.converged)
        r = dev.residual(H["go"], u, np.zeros(H["ndof"]))
        out["residual_at_solution"] = float(np.linalg.norm(r))
        off = dev.get_child_offsets()
        out["block_norms"] = [float(np.linalg.norm(u[off[k]:off[k + 1]])) for k in range(4)]
    dev.close()
    return out


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [40, 80]
    for n in sizes:
        print(json.dumps(run(n)), flush=True)
