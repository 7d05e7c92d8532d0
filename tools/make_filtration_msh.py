"""tools/make_filtration_msh.py — writes examples/filtration.msh, a Gmsh 2.1 ASCII triangle mesh in the layout of the reference's
test/stokesdarcy4.msh (which cannot travel to the GPU box): the square [0,100]^2, free flow (physical group 0) above y = 15, soil
(group 1) below with a low-permeability lens (group 2), boundary lines in physical group 7.  Deterministic (seed 4)."""
import os
import sys

import numpy as np


def build(nx=40, ny=40, X=100.0, Y=100.0, y_if=15.0, seed=4, jitter=0.2):
    rng = np.random.default_rng(seed)
    hx, hy = X / nx, Y / ny
    jline = int(round(y_if / hy))
    assert abs(jline * hy - y_if) < 1e-12
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    ix, iy = ix.ravel(), iy.ravel()
    xy = np.stack([ix * hx, iy * hy], axis=1).astype(float)
    inner_x = (ix > 0) & (ix < nx)
    inner_y = (iy > 0) & (iy < ny) & (iy != jline)
    xy[:, 0] += np.where(inner_x, rng.uniform(-jitter, jitter, xy.shape[0]) * hx, 0.0)
    xy[:, 1] += np.where(inner_y, rng.uniform(-jitter, jitter, xy.shape[0]) * hy, 0.0)
    vid = lambda i, j: j * (nx + 1) + i
    conn = []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            conn += [[a, b, d], [a, d, c]] if (i + j) % 2 == 0 else [[a, b, c], [b, d, c]]
    conn = np.array(conn)
    conn = conn[rng.permutation(conn.shape[0])]                     # no lattice order in the file
    cen = xy[conn].mean(axis=1)
    groups = np.where(cen[:, 1] > y_if, 0, 1)
    groups[(cen[:, 1] < y_if) & (np.abs(cen[:, 0] - 45.0) < 20.0) & (np.abs(cen[:, 1] - 7.5) < 3.0)] = 2
    return xy, conn, groups


def write(path, xy, conn, groups):
    edges = {}
    for e in range(conn.shape[0]):
        for a, b in ((0, 1), (0, 2), (1, 2)):
            edges.setdefault((min(conn[e, a], conn[e, b]), max(conn[e, a], conn[e, b])), []).append(e)
    lines = [k for k, v in sorted(edges.items()) if len(v) == 1]
    with open(path, "w") as fp:
        fp.write("$MeshFormat\n2.1 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % xy.shape[0])
        for i, p in enumerate(xy):
            fp.write("%d %.17g %.17g 0\n" % (i + 1, p[0], p[1]))
        fp.write("$EndNodes\n$Elements\n%d\n" % (len(lines) + conn.shape[0]))
        n = 1
        for a, b in lines:
            fp.write("%d 1 3 7 7 0 %d %d\n" % (n, a + 1, b + 1)); n += 1
        for e in range(conn.shape[0]):
            fp.write("%d 2 3 %d %d 0 %d %d %d\n" % (n, groups[e], groups[e], conn[e, 0] + 1, conn[e, 1] + 1, conn[e, 2] + 1)); n += 1
        fp.write("$EndElements\n")


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "examples", "filtration.msh")
    xy, conn, groups = build()
    write(out, xy, conn, groups)
    print("wrote", out, xy.shape[0], "vertices", conn.shape[0], "triangles", np.bincount(groups))
