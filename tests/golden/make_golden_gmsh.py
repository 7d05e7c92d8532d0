"""Golden sizes of the reference's own Gmsh fixtures (test/*.msh of /root/reference) as read by the independent Python reading
of the GmshReader rules (tests/um_problems.read_gmsh): vertices, triangles, cells per physical group, coordinate and
connectivity checksums.  Run in the build container (the fixtures cannot travel): `python tests/golden/make_golden_gmsh.py`.
tests/test_gmsh.py compares the C++ facade's reader against these numbers whenever /root/reference is present."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import um_problems as ump  # noqa: E402

REF = "/root/reference/test"
FILES = ["stokesdarcy2-fine.msh", "stokesdarcy4.msh", "stokesdarcy.msh", "stokesdarcy2.msh", "stokesdarcy3.msh", "stokesdarcy3-adaptive.msh",
         "stokesdarcy2fine.msh", "gmshtest.msh"]


def summary(xy, conn, groups):
    g, c = np.unique(groups, return_counts=True)
    return {"vertices": int(xy.shape[0]), "triangles": int(conn.shape[0]), "groups": {str(int(a)): int(b) for a, b in zip(g, c)},
            "sum_x": float(xy[:, 0].sum()), "sum_y": float(xy[:, 1].sum()), "sum_conn": int(conn.sum()),
            "weighted_conn": int((conn * np.array([1, 3, 7])).sum())}


if __name__ == "__main__":
    out = {}
    for f in FILES:
        p = os.path.join(REF, f)
        if os.path.exists(p):
            out[f] = summary(*ump.read_gmsh(p))
            print(f, out[f]["vertices"], out[f]["triangles"], out[f]["groups"])
    json.dump(out, open(os.path.join(HERE, "gmsh_fixture_sizes.json"), "w"), indent=1, sort_keys=True)
