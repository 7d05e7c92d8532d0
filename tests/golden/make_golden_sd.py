"""Generate the golden fixtures of the Stokes-Darcy path, tests/golden/sd/*.npz (run from the repo root:
`python tests/golden/make_golden_sd.py`).  NOT outputs of the reference (it cannot be built here and ships no expected values): outputs
of oracle/um_sd_oracle.py after it passed the pins of tests/test_oracle_sd.py, frozen so that a later change of the oracle cannot
silently move the target and so that the CUDA path is compared against committed numbers on the GPU box.  The mesh (coordinates,
connectivity, physical groups) is stored in the fixture; the parameters are those of tests/sd_problems.ini_params with the listed
overrides."""
import os
import sys

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sd_problems as sdp  # noqa: E402

CASES = {
    # name: (nx, ny, seed, ini_params overrides)
    "sd2d_6x4_shipped": (6, 4, 3, dict()),
    "sd2d_8x4_intended": (8, 4, 5, dict(quirk=0)),
    "sd2d_6x4_eps1e-8": (6, 4, 9, dict(epsilon=1e-8)),
}


def main():
    out = os.path.join(HERE, "sd")
    os.makedirs(out, exist_ok=True)
    for name, (nx, ny, seed, pk) in CASES.items():
        xy, conn, groups = sdp.filtration_mesh(nx, ny, seed=seed)
        o = sdp.oracle()
        H = sdp.setup(o, xy, conn, groups, params=sdp.ini_params(**pk))
        x0 = H["x0"]
        o.jacobian(H["go"], H["mat"], x0, overwrite=True)
        r = o.residual(H["go"], x0, np.zeros(H["ndof"]))
        rp, ci = o.matrix_get_pattern(H["mat"])
        val = o.matrix_get_values(H["mat"])
        A = sps.csr_matrix((val, ci, rp), shape=(H["ndof"],) * 2).tocsc()
        lu = spla.splu(A)
        z = lu.solve(r); z = z + lu.solve(r - A @ z)
        ptr, dofs = o.get_dofmap()
        np.savez_compressed(os.path.join(out, name + ".npz"), xy=xy, conn=conn, groups=groups,
                            pk_keys=np.array(sorted(pk)), pk_vals=np.array([float(pk[k]) for k in sorted(pk)]),
                            ndof=H["ndof"], ncon=H["ncon"], offsets=o.get_child_offsets(), cell_ptr=ptr, cell_dofs=dofs,
                            constrained=o.get_constraints(), rowptr=rp, colidx=ci, values=val, x0=x0, residual=r, solution_z=z)
        print(name, "ndof", H["ndof"], "nnz", ci.size, "ncon", H["ncon"])


if __name__ == "__main__":
    main()
