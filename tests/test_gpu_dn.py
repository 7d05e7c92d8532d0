"""Parity of the CUDA Neumann-Dirichlet coupling, the sub-block solver and the Dirichlet-Neumann iteration against the CPU oracle —
SURVEY §8 a10, BASELINE configs[0] (test/neumanndirichletcoupling.hh, test/testpoisson-dirichlet-neumann.cc with conforming spaces)."""
import importlib

import numpy as np
import pytest

import orc
import problems

pytestmark = pytest.mark.gpu
pkg = importlib.import_module("dune-multidomain_b200")
capi = pkg.capi

RTOL_ENTRY = 1e-12
RTOL_SOLUTION = 1e-10

CASES = {
    "dn2d_shipped_modes": dict(n=(32, 32)),                                                  # ini: left = Dirichlet, right = Neumann, alpha = 2
    "dn2d_swapped_modes": dict(n=(12, 10), left_mode=capi.ND_MODE_NEUMANN, right_mode=capi.ND_MODE_DIRICHLET, alpha=3.5),
    "dn2d_both": dict(n=(10, 8), left_mode=capi.ND_MODE_BOTH, right_mode=capi.ND_MODE_BOTH),
    "dn2d_q2q1": dict(n=(8, 6), fe=(capi.FE_Q2, capi.FE_Q1), right_mode=capi.ND_MODE_BOTH),
    "dn2d_q1q2": dict(n=(6, 8), fe=(capi.FE_Q1, capi.FE_Q2), left_mode=capi.ND_MODE_BOTH, right_mode=capi.ND_MODE_DIRICHLET),
    # the shipped instantiation of solve(): DG-Q2 spaces with ConvectionDiffusionDG on both sides (:968-971)
    "dn2d_dgq2_shipped": dict(n=(16, 16), fe=(capi.FE_DGQ2, capi.FE_DGQ2), dg=True),
    "dn2d_dgq1_both": dict(n=(10, 6), fe=(capi.FE_DGQ1, capi.FE_DGQ1), dg=True, left_mode=capi.ND_MODE_BOTH, right_mode=capi.ND_MODE_BOTH, alpha=3.0),
    "dn2d_dgq2_dgq1": dict(n=(6, 8), fe=(capi.FE_DGQ2, capi.FE_DGQ1), dg=True, left_mode=capi.ND_MODE_NEUMANN, right_mode=capi.ND_MODE_DIRICHLET),
}


def _pair(case):
    kw = CASES[case]
    dev, ora = pkg.Mdb(0), orc.Oracle()
    return dev, ora, problems.dirichlet_neumann(dev, **kw), problems.dirichlet_neumann(ora, **kw)


@pytest.fixture(params=sorted(CASES))
def pair(request):
    dev, ora, Pd, Po = _pair(request.param)
    yield dev, ora, Pd, Po
    dev.close(); ora.close()


def test_patterns_bit_exact(pair):
    dev, ora, Pd, Po = pair
    assert Pd.ndof == Po.ndof and Pd.ncon == Po.ncon
    for md, mo in ((Pd.mat, Po.mat), (Pd.lmat, Po.lmat), (Pd.rmat, Po.rmat)):
        rd, cd = dev.matrix_get_pattern(md); ro, co = ora.matrix_get_pattern(mo)
        assert np.array_equal(rd, ro) and np.array_equal(cd, co)


@pytest.mark.parametrize("which", ["lop", "rop", "full"])
def test_residual_and_jacobian_entries(pair, which):
    dev, ora, Pd, Po = pair
    x = np.random.default_rng(0).uniform(-1, 1, Po.ndof)
    god, goo = getattr(Pd, which), getattr(Po, which)
    rd, ro = np.zeros(Pd.ndof), np.zeros(Po.ndof)
    dev.residual(god, x, rd); ora.residual(goo, x, ro)
    assert np.abs(rd - ro).max() <= 10 * RTOL_ENTRY * np.abs(ro).max()
    dev.residual(god, x, rd, weight=-0.5); ora.residual(goo, x, ro, weight=-0.5)
    assert np.abs(rd - ro).max() <= 10 * RTOL_ENTRY * np.abs(ro).max()
    if which == "full":
        return
    md, mo = (Pd.lmat, Po.lmat) if which == "lop" else (Pd.rmat, Po.rmat)
    for overwrite, weight in ((True, 1.0), (False, 0.25)):
        dev.jacobian(god, md, x, weight=weight, overwrite=overwrite); ora.jacobian(goo, mo, x, weight=weight, overwrite=overwrite)
        vd, vo = dev.matrix_get_values(md), ora.matrix_get_values(mo)
        assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()


def test_3d_nd_coupling_entries():
    out = []
    for api in (pkg.Mdb(0), orc.Oracle()):
        api.mesh_structured((6, 4, 5)); api.subdomains_halfspace(1, 0.5, 0, 1)
        c0, c1 = api.add_space(capi.FE_Q1, 0), api.add_space(capi.FE_Q1, 1)
        pp = capi.PoissonParams(); pp.f = capi.Function(capi.FN_DN_SOURCE); pp.bc = capi.BCType(capi.BC_TESTPOISSON)
        pp.j = capi.Function(capi.FN_TESTPOISSON_J); pp.quad_order = 4
        l = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 1, [c0]); r = api.add_subproblem(capi.OP_POISSON, pp, capi.COND_EQUALITY, 2, [c1])
        q = capi.NDCouplingParams(); q.mode, q.method, q.weights, q.intorderadd, q.alpha = capi.ND_MODE_BOTH, capi.DG_METHOD_NIPG, capi.DG_WEIGHTS_OFF, 1, 1.5
        cp = api.add_coupling(capi.OP_ND_COUPLING, q, r, l)               # fired from the right subdomain: faces with side = 0
        nd = api.finalize(); api.constraints_assemble([l, r], [pp.bc, pp.bc])
        go = api.gridoperator_create([r], [cp]); mat = api.matrix_create([go])
        x = np.random.default_rng(4).uniform(-1, 1, nd)
        api.jacobian(go, mat, x, overwrite=True)
        rr = np.zeros(nd); api.residual(go, x, rr)
        out.append((api.matrix_get_pattern(mat), api.matrix_get_values(mat), rr))
    (pd_, vd, rd), (po, vo, ro) = out
    assert np.array_equal(pd_[0], po[0]) and np.array_equal(pd_[1], po[1])
    assert np.abs(vd - vo).max() <= RTOL_ENTRY * np.abs(vo).max()
    assert np.abs(rd - ro).max() <= 10 * RTOL_ENTRY * np.abs(ro).max()


def test_solve_block_matches_oracle(pair):
    """ISTL_SEQ_Subblock_Backend<SeqSSOR, BiCGSTABSolver> (:264-370): BiCGSTAB + SSOR(3) on the diagonal block; z outside the block
    is left untouched; iteration counts agree with the oracle's extracted-block solve."""
    dev, ora, Pd, Po = pair
    x = np.random.default_rng(1).uniform(-1, 1, Po.ndof)
    for (god, md, goo, mo, k) in ((Pd.lop, Pd.lmat, Po.lop, Po.lmat, 0), (Pd.rop, Pd.rmat, Po.rop, Po.rmat, 1)):
        dev.jacobian(god, md, x, overwrite=True); ora.jacobian(goo, mo, x, overwrite=True)
        r = np.zeros(Po.ndof); ora.residual(goo, x, r)
        zd, zo = np.full(Pd.ndof, 7.0), np.full(Po.ndof, 7.0)
        sd = dev.solve_block(md, Pd.children[k], r, zd, reduction=1e-12)
        so = ora.solve_block(mo, Po.children[k], r, zo, reduction=1e-12)
        assert sd.converged and so.converged and abs(sd.iterations - so.iterations) <= 1, (sd.iterations, so.iterations)
        assert np.abs(zd - zo).max() <= RTOL_SOLUTION * np.abs(zo).max()
        off = ora.get_child_offsets()
        outside = np.ones(Po.ndof, dtype=bool); outside[off[k]:off[k + 1]] = False
        assert np.all(zd[outside] == 7.0)


def test_solve_block_rejects_a_coupled_matrix():
    dev, ora, Pd, Po = _pair("dn2d_both")
    x = np.zeros(Pd.ndof)
    dev.jacobian(Pd.go, Pd.mat, x, overwrite=True)                      # the monolithic matrix has sn / ns blocks
    with pytest.raises(capi.MdbError):
        dev.solve_block(Pd.mat, Pd.children[0], x, np.zeros(Pd.ndof))
    dev.close(); ora.close()


@pytest.mark.parametrize("case", ["dn2d_shipped_modes", "dn2d_dgq2_shipped"])
def test_dn_iteration_matches_oracle(case):
    """The Dirichlet-Neumann loop (:721-822).  With tight inner solves the iterates agree to 1e-9; with the ini's loose inner
    reductions (1e-4) the outer iteration counts agree (+-1) and both reach the 1e-8 reduction of the full-operator residual."""
    dev, ora, Pd, Po = _pair(case)
    u0 = np.zeros(Po.ndof); ora.interpolate(Po.sps, Po.gfn, u0)
    tight = (1e-13, 1e-13, 0.2)
    hd, ho = [], []
    ud, _, _ = problems.dn_iteration(dev, Pd, u0.copy(), iterations=5, left_reduction=tight, right_reduction=tight, history=hd)
    uo, _, _ = problems.dn_iteration(ora, Po, u0.copy(), iterations=5, left_reduction=tight, right_reduction=tight, history=ho)
    assert np.abs(ud - uo).max() <= 1e-9 * np.abs(uo).max()
    assert np.allclose(hd, ho, rtol=1e-7)
    hd, ho = [], []
    ud, itd, rd = problems.dn_iteration(dev, Pd, u0.copy(), history=hd)
    uo, ito, ro = problems.dn_iteration(ora, Po, u0.copy(), history=ho)
    assert abs(itd - ito) <= 1 and hd[-1] / hd[0] <= 1e-8 * 40 and ho[-1] / ho[0] <= 1e-8 * 40
    assert np.abs(ud - uo).max() <= 1e-5 * np.abs(uo).max()
    dev.close(); ora.close()


@pytest.mark.parametrize("refine,ks", [(5, (1, 1)), (4, (2, 2)), (4, (1, 2)), (4, ("dg2", "dg2")), (4, ("dg1", "dg1"))])
def test_cxx_dn_driver_matches_oracle(refine, ks):
    """examples/testpoisson-dirichlet-neumann.cc with examples/poisson-dirichlet-neumann.ini (the reference's keys): DOF counts, the
    monolithic BiCGSTAB + SSOR(3) solve, the start residual of the full operator and the Dirichlet-Neumann sweep count / solution."""
    import os
    import re
    import subprocess
    import tempfile
    from test_capi import _build_example, ROOT
    n = (1 << refine,) * 2
    fe = tuple({1: capi.FE_Q1, 2: capi.FE_Q2, "dg1": capi.FE_DGQ1, "dg2": capi.FE_DGQ2}[k] for k in ks)      # ("dg2", "dg2") = as shipped
    ora = orc.Oracle()
    Po = problems.dirichlet_neumann(ora, n, fe=fe, dg=isinstance(ks[0], str))
    u0 = np.zeros(Po.ndof); ora.interpolate(Po.sps, Po.gfn, u0)
    ora.jacobian(Po.go, Po.mat, u0, overwrite=True)
    r = np.zeros(Po.ndof); ora.residual(Po.go, u0, r)
    z = np.zeros(Po.ndof); ora.solve(Po.mat, r, z, reduction=1e-13)
    umono = u0 - z
    rf = np.zeros(Po.ndof); ora.residual(Po.full, u0, rf)
    udn, its, _ = problems.dn_iteration(ora, Po, u0.copy())
    ini = open(os.path.join(ROOT, "examples", "poisson-dirichlet-neumann.ini")).read().replace("refine = 6", "refine = %d" % refine)
    with tempfile.TemporaryDirectory() as d:
        exe = _build_example(d, "testpoisson-dirichlet-neumann")
        open(os.path.join(d, "t.ini"), "w").write(ini)
        out = subprocess.run([exe, os.path.join(d, "t.ini"), str(ks[0]), str(ks[1])], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    m = re.search(r"(\d+) dof total, (\d+) dof constrained", out.stdout)
    assert m and int(m.group(1)) == Po.ndof and int(m.group(2)) == Po.ncon
    m = re.search(r"Monolithic solver: .* sum\(u\) = (\S+) \|u\|_2 = (\S+)", out.stdout)
    assert abs(float(m.group(2)) - np.linalg.norm(umono)) <= 1e-6 * np.linalg.norm(umono)       # monolithic reduction 1e-8
    m = re.search(r"Start residual: (\S+)", out.stdout)
    assert abs(float(m.group(1)) - np.linalg.norm(rf)) <= 1e-11 * np.linalg.norm(rf)
    m = re.search(r"Dirichlet-Neumann solver: (\d+) sweeps, (\d+) inner iterations, sum\(u\) = (\S+) \|u\|_2 = (\S+)", out.stdout)
    assert m and abs(int(m.group(1)) - its) <= 1
    assert abs(float(m.group(4)) - np.linalg.norm(udn)) <= 1e-5 * np.linalg.norm(udn)
