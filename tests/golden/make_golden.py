"""Generate the golden fixtures of tests/golden/ (run from the repo root: `python tests/golden/make_golden.py`).

The reference cannot be built or imported in this image and ships no expected values, so these vectors
are NOT outputs of the reference.  They are outputs of the CPU oracle (oracle/oracle.cpp) after it passed the
known-answer pins of tests/test_oracle.py, frozen so that (a) a later change of the oracle cannot silently move the
target and (b) the CUDA path is compared against committed numbers on the GPU box, where /root/reference does not exist.
Each fixture also stores an INDEPENDENT NumPy/SciPy construction where one exists cheaply:
  * kron_volume: the two-subdomain Q1 Laplace matrix assembled from 1-D stiffness/mass matrices by Kronecker products
    (no quadrature loop, no element traversal) — must agree with the oracle's volume rows to 1e-13.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orc  # noqa: E402
import problems  # noqa: E402

capi = orc.capi

CASES = {
    # name: (cells, kwargs of problems.testpoisson)
    "p2d_16x16_cvcf2": ((16, 16), dict(intensity=2.0)),                                  # testpoisson "4 2.0" with Q1/Q1
    "p3d_4x6x5_cvcf2": ((4, 6, 5), dict(intensity=2.0)),
    "p3d_5x4x4_propflow": ((5, 4, 4), dict(intensity=3.0, coupling=capi.OP_PROPFLOW_COUPLING)),
    "p2d_9x8_split0": ((9, 8), dict(intensity=10.0, split_axis=0)),
    # the reference's shipped driver as it stands: Q1 on the left subdomain, Q2 on the right (test/testpoisson.cc:168-198), "4 2.0"
    "p2d_16x16_q1q2": ((16, 16), dict(intensity=2.0, fe=(capi.FE_Q1, capi.FE_Q2))),
    "p3d_3x4x3_q2q1": ((3, 4, 3), dict(intensity=2.0, fe=(capi.FE_Q2, capi.FE_Q1))),
    # test/testmortarpoisson.cc as shipped ("5 1.0": 32x32 cells, Q1 | Q1 | Q1 mortar space, 1155 DOFs) and a 3-D Q1|Q2 variant
    "p2d_32x32_mortar": ((32, 32), dict(problem="mortar", gamma_0=1.0)),
    "p3d_4x3x3_mortar_q1q2": ((4, 3, 3), dict(problem="mortar", gamma_0=2.5, fe=(capi.FE_Q1, capi.FE_Q2))),
    # test/testpowermortarpoisson.cc (PowerCouplingGridFunctionSpace<CouplingGFS,2,VBE>) in 3-D, where its matrix is regular
    "p3d_6x5x4_powermortar": ((6, 5, 4), dict(problem="mortar", gamma_0=1.0, ncomp=2)),
    # test/testpoisson-dirichlet-neumann.cc as shipped (DG-Q2 | DG-Q2, ConvectionDiffusionDG(SIPG, weightsOn, 2), CVCF(4, 10)) at
    # mesh.refine = 3, and a 3-D DG-Q1 variant
    "dg2d_8x8_q2_shipped": ((8, 8), dict(problem="dg", k=2, scaling=10.0)),
    "dg3d_4x3x2_q1": ((4, 3, 2), dict(problem="dg", k=1, scaling=3.0)),
}


def kron_volume(n, split_axis):
    """Independent construction of the volume part: on each subdomain box the Q1 Laplace matrix is
    sum_d  K_d (x) M_other...  with the 1-D element matrices K = 1/h [[1,-1],[-1,1]], M = h/6 [[2,1],[1,2]]."""
    dim = len(n)

    def k1(m, h):
        a = sp.lil_matrix((m + 1, m + 1))
        for e in range(m):
            a[e, e] += 1 / h; a[e + 1, e + 1] += 1 / h; a[e, e + 1] -= 1 / h; a[e + 1, e] -= 1 / h
        return a.tocsr()

    def m1(m, h):
        a = sp.lil_matrix((m + 1, m + 1))
        for e in range(m):
            a[e, e] += h / 3; a[e + 1, e + 1] += h / 3; a[e, e + 1] += h / 6; a[e + 1, e] += h / 6
        return a.tocsr()

    blocks = []
    for s in range(2):
        cells = list(n)
        cells[split_axis] = n[split_axis] // 2 if s == 0 else n[split_axis] - n[split_axis] // 2
        hs = [1.0 / n[d] for d in range(dim)]
        total = None
        for d in range(dim):
            term = None
            for a in reversed(range(dim)):              # axis 0 fastest => last Kronecker factor
                f = k1(cells[a], hs[a]) if a == d else m1(cells[a], hs[a])
                term = f if term is None else sp.kron(term, f, format="csr")
            total = term if total is None else total + term
        blocks.append(total)
    return sp.block_diag(blocks, format="csr")


def main():
    for name, (n, kw) in CASES.items():
        if os.path.exists(os.path.join(HERE, name + ".npz")) and "--all" not in sys.argv:
            continue                                    # frozen fixtures are not regenerated unless asked for
        o = orc.Oracle()
        kw = dict(kw)
        problem = kw.pop("problem", "poisson")
        if problem == "dg": P = problems.dirichlet_neumann_dg(o, n, **kw)
        else: P = problems.testmortarpoisson(o, n, **kw) if problem == "mortar" else problems.testpoisson(o, n, **kw)
        ptr, dofs = o.get_dofmap()
        rp, ci = o.matrix_get_pattern(P.mat)
        u = np.zeros(P.ndof)
        o.interpolate(P.sps, P.gfn, u)
        o.jacobian(P.go, P.mat, u, overwrite=True)
        vals = o.matrix_get_values(P.mat)
        r = np.zeros(P.ndof)
        o.residual(P.go, u, r)
        xr = np.random.default_rng(0).uniform(-1, 1, P.ndof)
        o.jacobian(P.go, P.mat, xr, overwrite=True)
        vals_rand = o.matrix_get_values(P.mat)
        rr = np.zeros(P.ndof)
        o.residual(P.go, xr, rr)
        o.jacobian(P.go, P.mat, u, overwrite=True)
        z = np.zeros(P.ndof)
        o.solve(P.mat, r, z, reduction=1e-13)
        # independent check of the volume rows: rows that are neither constrained nor touched by the coupling
        split = kw.get("split_axis", 0 if problem in ("mortar", "dg") else 1)
        fe = kw.get("fe", (capi.FE_Q1, capi.FE_Q1))
        if n[split] % 2 == 0 and fe == (capi.FE_Q1, capi.FE_Q1) and problem == "poisson":
            K = kron_volume(n, split)
            A = sp.csr_matrix((vals, ci, rp), shape=(P.ndof, P.ndof))
            con = o.get_constraints().astype(bool)
            off = o.get_child_offsets()
            diff = (A - K).tocsr()
            cross = np.zeros(P.ndof, dtype=bool)          # rows with entries in the other child's block = interface rows
            for row in range(P.ndof):
                c = ci[rp[row]:rp[row + 1]]
                cross[row] = ((c < off[1]) != (row < off[1])).any()
            free = ~con & ~cross
            err = abs(diff[free]).max()
            assert err < 1e-13, (name, err)
            print("%s: Kronecker volume check on %d rows, max diff %.2e" % (name, int(free.sum()), err))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), problem=problem, ncomp=kw.get("ncomp", 1), k=kw.get("k", 0), scaling=kw.get("scaling", 0.0),
                            n=np.array(n), split_axis=split, intensity=kw.get("intensity", 2.0),
                            coupling=kw.get("coupling", capi.OP_MORTAR_POISSON_COUPLING if problem == "mortar" else capi.OP_CVCF_COUPLING),
                            gamma_0=kw.get("gamma_0", 0.0), child_offsets=o.get_child_offsets(), fe=np.array(fe), ndof=P.ndof, ncon=P.ncon,
                            cell_ptr=ptr, cell_dofs=dofs, constrained=o.get_constraints(), rowptr=rp, colidx=ci, u=u,
                            jac_at_u=vals, res_at_u=r, x_rand=xr, jac_at_rand=vals_rand, res_at_rand=rr, solution_z=z)
        print("%s: ndof %d ncon %d nnz %d" % (name, P.ndof, P.ncon, len(ci)))


if __name__ == "__main__":
    main()
