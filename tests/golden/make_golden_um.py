"""Generate the golden fixtures of the unstructured-mesh path, tests/golden/um/*.npz (run from the repo root:
`python tests/golden/make_golden_um.py`).  Like tests/golden/make_golden.py these are NOT outputs of the reference (it cannot be
built here and ships no expected values): they are outputs of oracle/um_oracle.py after it passed the pins of
tests/test_oracle_um.py, frozen so that a later change of the oracle cannot silently move the target and so that the CUDA path is
compared against committed numbers.  The mesh itself (coordinates, connectivity, subdomain marks) is stored in the fixture."""
import os
import sys

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import um_problems as ump  # noqa: E402

CASES = {
    # name: (nx, ny, seed, split axis, kwargs of um_problems.setup)
    "um2d_6x5_propflow_fd": (6, 5, 3, 0, dict()),
    "um2d_9x7_split_y_mass": (9, 7, 8, 1, dict(with_mass=True, k=4.0)),
    "um2d_5x5_shared_space": (5, 5, 1, 0, dict(shared_space=True)),
    "um2d_8x6_cvcf": (8, 6, 4, 0, dict(cvcf=True, k=2.0)),
}


def main():
    out = os.path.join(HERE, "um")
    os.makedirs(out, exist_ok=True)
    for name, (nx, ny, seed, axis, kw) in CASES.items():
        xy, conn = ump.square_mesh(nx, ny, seed=seed)
        sd = ump.halves(xy, conn, axis=axis)
        o = ump.oracle()
        H = ump.setup(o, xy, conn, sd, **kw)
        x0 = H["x0"]
        o.jacobian(H["go"], H["mat"], x0, overwrite=True)
        r = o.residual(H["go"], x0, np.zeros(H["ndof"]))
        rp, ci = o.matrix_get_pattern(H["mat"])
        val = o.matrix_get_values(H["mat"])
        u = x0 - spla.spsolve(sps.csr_matrix((val, ci, rp), shape=(H["ndof"],) * 2).tocsc(), r)
        ptr, dofs = o.get_dofmap()
        np.savez_compressed(os.path.join(out, name + ".npz"), xy=xy, conn=conn, sd=sd, axis=axis,
                            kw_keys=np.array(sorted(kw)), kw_vals=np.array([float(kw[k]) for k in sorted(kw)]),
                            ndof=H["ndof"], ncon=H["ncon"], offsets=o.get_child_offsets(), cell_ptr=ptr, cell_dofs=dofs,
                            constrained=o.get_constraints(), rowptr=rp, colidx=ci, values=val, x0=x0, residual=r, solution=u)
        print(name, "ndof", H["ndof"], "nnz", ci.size, "ncon", H["ncon"])


if __name__ == "__main__":
    main()
