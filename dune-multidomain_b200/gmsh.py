"""Gmsh 2.x ASCII triangle meshes for the bench and test plumbing: the reading rules of the facade's GmshReader
(include/mdb200/multidomain.hh; [UPSTREAM] dune-grid gmshreader.hh as the reference calls it, test/stokesdarcy2.cc:575-583) and
regular refinement, vectorised.  Only nodes used by a line or a triangle become vertices, numbered in order of first appearance in
$Elements; triangles keep the file order; the first tag is the physical group."""
import numpy as np


def read(path):
    """Returns (xy [nv, 2], conn [ne, 3], groups [ne])."""
    with open(path) as fp:
        tok = fp.read().split()
    i = tok.index("$Nodes") + 1
    nn = int(tok[i]); i += 1
    raw = np.array(tok[i:i + 4 * nn], dtype=np.float64).reshape(nn, 4)
    ids = raw[:, 0].astype(np.int64)
    coords = np.zeros((int(ids.max()) + 1, 2))
    coords[ids] = raw[:, 1:3]
    i = tok.index("$Elements") + 1
    nel = int(tok[i]); i += 1
    order, tris, groups = [], [], []
    for _ in range(nel):
        etype, ntags = int(tok[i + 1]), int(tok[i + 2])
        tags = tok[i + 3:i + 3 + ntags]
        nnodes = {1: 2, 2: 3, 15: 1}.get(etype)
        if nnodes is None:
            raise ValueError("gmsh element type %d is not supported (lines, triangles, points)" % etype)
        nodes = [int(v) for v in tok[i + 3 + ntags:i + 3 + ntags + nnodes]]
        i += 3 + ntags + nnodes
        if etype == 15:
            continue
        order += nodes
        if etype == 2:
            tris.append(nodes); groups.append(int(tags[0]) if ntags else 0)
    order = np.array(order, dtype=np.int64)
    _, first = np.unique(order, return_index=True)
    used = order[np.sort(first)]                                  # order of first appearance
    new = -np.ones(coords.shape[0], dtype=np.int64)
    new[used] = np.arange(used.size)
    return np.ascontiguousarray(coords[used]), new[np.array(tris, dtype=np.int64)], np.array(groups, dtype=np.int32)


def refine(xy, conn, groups, levels=1):
    """Regular ("red") refinement as GmshMesh::globalRefine of the facade: children (v0,m01,m02), (m01,v1,m12), (m02,m12,v2),
    (m01,m12,m02) replace their father in place and inherit its group.  The midpoints are numbered by ascending edge key here (the
    facade numbers them in order of creation): same mesh, another vertex order — bench meshes only."""
    for _ in range(levels):
        nv = xy.shape[0]
        a = np.concatenate([conn[:, 0], conn[:, 0], conn[:, 1]])
        b = np.concatenate([conn[:, 1], conn[:, 2], conn[:, 2]])
        key = np.minimum(a, b) * nv + np.maximum(a, b)
        uk, inv = np.unique(key, return_inverse=True)
        mid = 0.5 * (xy[uk // nv] + xy[uk % nv])
        ne = conn.shape[0]
        m01, m02, m12 = nv + inv[:ne], nv + inv[ne:2 * ne], nv + inv[2 * ne:]
        v0, v1, v2 = conn[:, 0], conn[:, 1], conn[:, 2]
        child = np.stack([np.stack([v0, m01, m02], 1), np.stack([m01, v1, m12], 1), np.stack([m02, m12, v2], 1), np.stack([m01, m12, m02], 1)], axis=1)
        conn = child.reshape(4 * ne, 3)
        groups = np.repeat(groups, 4)
        xy = np.concatenate([xy, mid])
    return np.ascontiguousarray(xy), np.ascontiguousarray(conn), np.ascontiguousarray(groups)
