"""Stokes-Darcy test problems (SURVEY §8 a5 + a12), driven identically through the product binding (capi.Mdb) and the Python oracle
(oracle/um_sd_oracle.SDOracle): the spaces, subproblems, coupling and data of test/stokesdarcy2.cc + test/topbottomfiltration.ini on
generated triangle meshes (free flow above, two soil layers below, like test/stokesdarcy4.msh: physical groups 0 | 1 | 2)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
capi = importlib.import_module("dune-multidomain_b200").capi

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import um_sd_oracle  # noqa: E402  (test infrastructure)
import um_problems as ump  # noqa: E402


def oracle():
    return um_sd_oracle.SDOracle()


def filtration_mesh(nx, ny, seed=0, X=100.0, Y=20.0, y_if=15.0, jitter=0.15, shuffle=True):
    """[0,X] x [0,Y] in nx x ny squares cut into triangles; ny must make y_if a grid line.  Returns xy, conn, groups:
    group 0 = free flow (y > y_if), 1 = soil, 2 = low-permeability lens (a box inside the soil)."""
    xy, conn = ump.square_mesh(nx, ny, seed=seed, jitter=0.0, shuffle=False, lower=(0.0, 0.0), upper=(X, Y))
    rng = np.random.default_rng(seed)
    hx, hy = X / nx, Y / ny
    jline = int(round(y_if / hy))
    assert abs(jline * hy - y_if) < 1e-12
    if jitter > 0:
        ix, iy = np.arange(xy.shape[0]) % (nx + 1), np.arange(xy.shape[0]) // (nx + 1)
        inner_x = (ix > 0) & (ix < nx)
        inner_y = (iy > 0) & (iy < ny) & (iy != jline)                 # the interface stays straight
        xy[:, 0] += np.where(inner_x, rng.uniform(-jitter, jitter, xy.shape[0]) * hx, 0.0)
        xy[:, 1] += np.where(inner_y, rng.uniform(-jitter, jitter, xy.shape[0]) * hy, 0.0)
    if shuffle:
        pv = rng.permutation(xy.shape[0])
        xy2 = np.empty_like(xy); xy2[pv] = xy
        conn = pv[conn]
        conn = conn[rng.permutation(conn.shape[0])]
        for e in range(conn.shape[0]):
            conn[e] = np.roll(conn[e], rng.integers(0, 3))
            if rng.integers(0, 2):
                conn[e] = conn[e][::-1]
        xy = xy2
    c = xy[conn].mean(axis=1)
    groups = np.where(c[:, 1] > y_if, 0, 1).astype(np.int32)
    lens = (c[:, 1] < y_if) & (np.abs(c[:, 0] - 0.45 * X) < 0.2 * X) & (np.abs(c[:, 1] - 0.5 * y_if) < 0.2 * y_if)
    groups[lens] = 2
    return np.ascontiguousarray(xy), np.ascontiguousarray(conn), groups


def ini_params(X=100.0, y_dirichlet=13.0, epsilon=1e-2, quirk=1, stokes_bc=capi.BC_ALL_NEUMANN, darcy_bc=None, darcy_k=3):
    """test/topbottomfiltration.ini:1-20 through the parameter classes of test/stokesdarcy2.cc."""
    th = capi.TaylorHoodParams()
    th.mu, th.rho = 1e-3, 1e3
    th.f[0], th.f[1] = 0.0, -9.81                                         # NavierStokesParameters::f = (0, -gravity)  :239-249
    th.bc = capi.BCType(stokes_bc)                                        # StokesBoundaryType: StressNeumann everywhere  :139-140
    th.j = capi.Function(capi.FN_PRESSURE_DROP, 5e-8, 4e-8, X, 50.0, 100.0)
    th.quad_order, th.full_tensor = 4, 1
    da = capi.DarcyParams()
    da.kabs, da.kabs_low, da.high_group = 1.16e-5, 1.16e-10, 1
    da.bc = capi.BCType(capi.BC_DARCY_RIGHT, X, y_dirichlet) if darcy_bc is None else darcy_bc
    da.g = capi.Function(capi.FN_DARCY_G, 14.8, 11.0, 15.0)
    da.j = capi.Function(capi.FN_DARCY_J, 1e-8, X)
    da.method, da.weights, da.intorderadd, da.alpha = capi.DG_METHOD_SIPG, capi.DG_WEIGHTS_ON, 0, 10.0
    cp = capi.StokesDarcyParams()
    cp.gravity, cp.alpha, cp.viscosity, cp.porosity, cp.gamma, cp.density, cp.kabs, cp.epsilon = 9.81, 1.0, 1e-3, 0.01, 1.0, 1e3, 1.16e-5, epsilon
    cp.quirk = quirk
    return th, da, cp


def setup(lib, xy, conn, groups, *, jac_mode=capi.JAC_FD_BITMATCH, params=None, darcy_fe=capi.FE_OPB3, coupling=True,
          stokes=True, darcy=True):
    """test/stokesdarcy2.cc:575-760: spaces [v_0 | v_1 | p | phi], Stokes subproblem over THREE children, Darcy subproblem, coupling."""
    th, da, cpp = ini_params() if params is None else params
    lib.mesh_unstructured(xy, conn)
    lib.subdomains(np.where(groups > 0, 2, 1).astype(np.uint8))           # physical group > 0 -> subdomain 1 (Darcy)  :596-601
    lib.cell_groups(groups)
    H = dict(params=(th, da, cpp))
    sps = []
    if stokes:
        v0, v1, p = lib.add_space(capi.FE_P2, 0), lib.add_space(capi.FE_P2, 0), lib.add_space(capi.FE_P1, 0)
        H["sp_stokes"] = lib.add_subproblem(capi.OP_TAYLORHOOD_STOKES, th, capi.COND_EQUALITY, 1, [v0, v1, p])
        sps.append(H["sp_stokes"])
    if darcy:
        phi = lib.add_space(darcy_fe, 1)
        H["sp_darcy"] = lib.add_subproblem(capi.OP_DARCY_DG, da, capi.COND_EQUALITY, 2, [phi])
        sps.append(H["sp_darcy"])
    cps = []
    if coupling and stokes and darcy:
        cps = [lib.add_coupling(capi.OP_STOKESDARCY_COUPLING, cpp, H["sp_stokes"], H["sp_darcy"], jac_mode)]
    H["ndof"] = lib.finalize()
    H["ncon"] = lib.constraints_assemble([H["sp_stokes"]], [th.bc]) if stokes else 0
    H["go"] = lib.gridoperator_create(sps, cps)
    H["mat"] = lib.matrix_create([H["go"]])
    rng = np.random.default_rng(1234)
    H["x0"] = rng.uniform(-1.0, 1.0, H["ndof"])
    return H


# ---- committed golden vectors (tests/golden/sd, written by tests/golden/make_golden_sd.py) -----------------------------------
def golden_cases():
    import glob
    return sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "sd", "*.npz")))


def check_golden(lib, path, solve=None):
    """DOF map, constraints, pattern bit-exact; entries 1e-12 of the largest; the Newton step (direct solve) 1e-10 per block."""
    g = np.load(path)
    pk = {str(k): float(v) for k, v in zip(g["pk_keys"], g["pk_vals"])}
    if "quirk" in pk:
        pk["quirk"] = int(pk["quirk"])
    H = setup(lib, g["xy"], g["conn"], g["groups"], params=ini_params(**pk))
    assert H["ndof"] == int(g["ndof"]) and H["ncon"] == int(g["ncon"])
    assert np.array_equal(lib.get_child_offsets(), g["offsets"])
    ptr, dofs = lib.get_dofmap()
    assert np.array_equal(ptr, g["cell_ptr"]) and np.array_equal(dofs, g["cell_dofs"])
    assert np.array_equal(lib.get_constraints(), g["constrained"])
    rp, ci = lib.matrix_get_pattern(H["mat"])
    assert np.array_equal(rp, g["rowptr"]) and np.array_equal(ci, g["colidx"])
    x0 = np.ascontiguousarray(g["x0"])
    assert np.array_equal(x0, H["x0"])
    lib.jacobian(H["go"], H["mat"], x0, overwrite=True)
    v = lib.matrix_get_values(H["mat"])
    assert np.abs(v - g["values"]).max() <= 1e-12 * np.abs(g["values"]).max()
    r = lib.residual(H["go"], x0, np.zeros(H["ndof"]))
    assert np.abs(r - g["residual"]).max() <= 1e-12 * np.abs(g["residual"]).max()
    if solve is not None:
        z = solve(H, np.ascontiguousarray(g["residual"]))
        off = g["offsets"]
        for a in range(len(off) - 1):
            blk = slice(int(off[a]), int(off[a + 1]))
            assert np.abs(z[blk] - g["solution_z"][blk]).max() <= 1e-10 * np.abs(g["solution_z"][blk]).max(), a
