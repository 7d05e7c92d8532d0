"""oracle/um_oracle.py — TEST INFRASTRUCTURE ONLY (parity oracle for the unstructured-mesh path).

PARITY UNPINNED at the upstream boundary: the reference (smuething/dune-multidomain) cannot be built here and
its tests hold no expected values.  This file restates, in plain Python loops for SMALL meshes, what the reference's assembler
does on an unstructured conforming triangle mesh (the grid class of test/stokesdarcy2.cc:575-604) with P1 spaces
(PkLocalFiniteElementMap<..., 1>, test/stokesdarcy2.cc:614), cell by cell in index order exactly like
GlobalAssembler::assemble (globalassembler.hh:43-298):

  numbering     multidomaingridfunctionspace.hh:308-369; children concatenated lexicographically, a P1 child on subdomain s
                numbers the subdomain's vertices by ascending host vertex index [UPSTREAM (ii), (iii)]
  local spaces  multidomainlocalfunctionspace.hh:348-358 (children present on the cell, concatenated)
  participants  subproblem.hh:84-87, coupling.hh:78-83, subdomainset.hh:102-131
  pattern       patternassemblerengine.hh:142-187 (FullVolumePattern), couplingutilities.hh:35-47 (sn / ns links per coupling face)
  residual      residualassemblerengine.hh:190-567, functors residualassemblerfunctors.hh:16-257 (interior face whose outside cell is
                not in the subproblem -> the operator's boundary methods, which answer "no condition" there)
  Jacobian      jacobianassemblerengine.hh:195-504; coupling Jacobian by NumericalJacobianCoupling couplingutilities.hh:75-135
  operators     [UPSTREAM] PDELab::Poisson = test/instationarypoissonoperator.hh:30-168, PDELab::L2 = test/simpletimeoperator.hh:29-101,
                ProportionalFlowCoupling test/proportionalflowcoupling.hh:27-86, ContinuousValueContinuousFlowCoupling :113-233
  constraints   constraints.hh:312-558, constraintsfunctors.hh:75-103 ([UPSTREAM] ConformingDirichletConstraints: predicate at the
                face centre, the face's vertices are constrained)
  interpolation interpolate.hh:35-73

[UPSTREAM] assumptions: DUNE simplex reference element (vertices (0,0),(1,0),(0,1); face 0 = {0,1}, 1 = {0,2}, 2 = {1,2}); P1
shape functions 1-x-y, x, y; Gauss-Legendre line rules; symmetric triangle rules (1, 3 and 6 points) for orders <= 1, 2, <= 4.
Pinned by tests/test_oracle_um.py (cotangent formula, SciPy adjacency, patch test, polynomial exactness, FD vs exact, SciPy solve).

Only tests/ and __graft_entry__.smoke() may import this.  The method names follow the ctypes binding of the product
(dune-multidomain_b200/capi.py) so that one problem-setup function drives both.
"""
import math

import numpy as np

FACE_V = ((0, 1), (0, 2), (1, 2))
REF = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))
OP_POISSON, OP_L2, OP_CVCF, OP_PROPFLOW = 1, 2, 101, 102
JAC_FD, JAC_ANALYTIC = 0, 1


def eval_function(f, x, t=0.0):
    """include/mdb200.h MDB_FN_* (problem data of the reference's drivers, e.g. test/testpoisson.cc:32-116)."""
    k, p = f.kind, f.p
    if k == 0:
        return 0.0
    if k == 1:
        return p[0]
    if k == 2:
        n2 = 0.0
        for d in range(2):
            c = p[1 + d] - x[d]
            n2 += c * c
        return math.exp(-(p[0] * n2))
    if k == 3:
        for d in range(2):
            if not (x[d] > p[1] and x[d] < p[2]):
                return 0.0
        return p[0]
    if k == 4:
        if x[1] < 1e-6 or x[1] > 1.0 - 1e-6:
            return 0.0
        if x[0] > 1.0 - 1e-6 and x[1] > 0.5 + 1e-6:
            return -5.0
        return 0.0
    if k == 5:                                            # test/testinstationarypoisson.cc:94-100 family (time dependent Dirichlet data)
        n2 = 0.0
        for d in range(2):
            c = 0.5 - x[d]
            n2 += c * c
        return p[0] * math.sin(p[1] * t) * math.exp(-n2)
    if k == 6:
        v, n2 = p[0], 0.0
        for d in range(2):
            v += p[1 + d] * x[d]
            n2 += x[d] * x[d]
        return v + p[4] * n2
    raise ValueError("function kind %d not restated for unstructured meshes" % k)


def eval_bctype(b, x, is_boundary):
    """0 Dirichlet, 1 Neumann, 2 none (test/testpoisson.cc:42-88)."""
    if b.kind == 0:
        return 0 if is_boundary else 2
    if b.kind == 1:
        return 1 if is_boundary else 2
    if b.kind == 2:
        if not is_boundary:
            return 2
        if x[1] < 1e-6 or x[1] > 1.0 - 1e-6:
            return 1
        if x[0] > 1.0 - 1e-6 and x[1] > 0.5 + 1e-6:
            return 1
        return 0
    return 2


def cond_applies(kind, mask, sd):
    return sd == mask if kind == 0 else (sd & mask) == mask


def tri_rule(order):
    if order <= 1:
        return [((1.0 / 3.0, 1.0 / 3.0), 0.5)]
    if order == 2:
        a, b = 1.0 / 6.0, 2.0 / 3.0
        return [((a, a), 1.0 / 6.0), ((b, a), 1.0 / 6.0), ((a, b), 1.0 / 6.0)]
    if order > 4:
        raise ValueError("simplex quadrature rules up to order 4")
    a, wa = 0.445948490915965, 0.5 * 0.223381589678011
    b, wb = 0.091576213509771, 0.5 * 0.109951743655322
    a2, b2 = 1.0 - 2.0 * a, 1.0 - 2.0 * b
    return [((a, a), wa), ((a2, a), wa), ((a, a2), wa), ((b, b), wb), ((b2, b), wb), ((b, b2), wb)]


def line_rule(order):
    m = (order + 2) // 2
    if m <= 1:
        return [(0.5, 1.0)]
    if m == 2:
        a = 0.5 / math.sqrt(3.0)
        return [(0.5 - a, 0.5), (0.5 + a, 0.5)]
    if m == 3:
        a = 0.5 * math.sqrt(3.0 / 5.0)
        return [(0.5 - a, 5.0 / 18.0), (0.5, 8.0 / 18.0), (0.5 + a, 5.0 / 18.0)]
    raise ValueError("line rule order too high")


def p1_basis(x, y):
    return (1.0 - x - y, x, y)


class _Sp:
    pass


class UMOracle:
    """CPU restatement; same call shapes as dune-multidomain_b200.capi.Mdb for the unstructured path."""

    def __init__(self):
        self.children, self.sps, self.cps, self.gos, self.mats = [], [], [], [], []
        self.time = 0.0

    # -- mesh --
    def mesh_unstructured(self, xy, conn):
        self.xy = np.asarray(xy, dtype=np.float64)
        self.conn = np.asarray(conn, dtype=np.int64)
        self.nv, self.ne = self.xy.shape[0], self.conn.shape[0]
        self.ncells = self.ne
        edge = {}
        self.nbr = -np.ones((self.ne, 3), dtype=np.int64)
        for e in range(self.ne):
            for f in range(3):
                a, b = self.conn[e, FACE_V[f][0]], self.conn[e, FACE_V[f][1]]
                key = (min(a, b), max(a, b))
                if key in edge:
                    e2, f2 = edge[key]
                    self.nbr[e, f], self.nbr[e2, f2] = e2, e
                else:
                    edge[key] = (e, f)

    def subdomains(self, bits):
        self.sd = np.asarray(bits, dtype=np.uint8).astype(np.int64)

    def add_space(self, fe_kind, subdomain):
        assert fe_kind == 21
        self.children.append(dict(subdomain=subdomain))
        return len(self.children) - 1

    def add_subproblem(self, op_id, params, cond_kind, cond_mask, children):
        s = _Sp()
        s.op_id, s.P, s.kind, s.mask, s.child = op_id, params, cond_kind, cond_mask, int(children[0])
        s.quad_order = params.quad_order
        self.sps.append(s)
        return len(self.sps) - 1

    def add_coupling(self, op_id, params, local_sp, remote_sp, jac_mode=JAC_FD):
        assert op_id in (OP_PROPFLOW, OP_CVCF)
        order = params.quad_order if op_id == OP_CVCF else 4          # ProportionalFlowCoupling: intorder fixed to 4 (:43)
        self.cps.append(dict(op=op_id, k=params.intensity, order=order, local=local_sp, remote=remote_sp, jac_mode=jac_mode))
        return len(self.cps) - 1

    def _present(self, k, e):
        s = self.children[k]["subdomain"]
        return s < 0 or ((self.sd[e] >> s) & 1) == 1

    def _applies(self, s, e):
        return cond_applies(self.sps[s].kind, self.sps[s].mask, int(self.sd[e]))

    # -- numbering --
    def finalize(self):
        off = 0
        for k, ch in enumerate(self.children):
            flag = np.zeros(self.nv, dtype=bool)
            for e in range(self.ne):
                if self._present(k, e):
                    flag[self.conn[e]] = True
            v2d = -np.ones(self.nv, dtype=np.int64)
            v2d[flag] = np.arange(int(flag.sum()))          # ascending host vertex index
            ch["v2d"], ch["size"], ch["offset"] = v2d, int(flag.sum()), off
            off += ch["size"]
        self.ndof = off
        self.constrained = np.zeros(self.ndof, dtype=np.uint8)
        return self.ndof

    def _dofs(self, k, e):
        ch = self.children[k]
        return [ch["offset"] + int(ch["v2d"][v]) for v in self.conn[e]]

    def get_dofmap(self):
        ptr, dofs = [0], []
        for e in range(self.ne):
            for k in range(len(self.children)):
                if self._present(k, e):
                    dofs += self._dofs(k, e)
            ptr.append(len(dofs))
        return np.array(ptr, dtype=np.int64), np.array(dofs, dtype=np.int64)

    def get_child_offsets(self):
        return np.array([ch["offset"] for ch in self.children] + [self.ndof], dtype=np.int64)

    # -- geometry --
    def _tri(self, e):
        p = [self.xy[v] for v in self.conn[e]]
        ax, ay = p[1][0] - p[0][0], p[1][1] - p[0][1]
        bx, by = p[2][0] - p[0][0], p[2][1] - p[0][1]
        det = ax * by - ay * bx
        inv = 1.0 / det            # [UPSTREAM] dune-common FMatrixHelp::invertMatrix (2x2) multiplies by the reciprocal determinant
        t = ((by * inv, -(ay * inv)), (-(bx * inv), ax * inv))
        js = ((-1.0, -1.0), (1.0, 0.0), (0.0, 1.0))
        gp = []
        for i in range(3):
            g = []
            for d in range(2):
                s = 0.0
                s += t[d][0] * js[i][0]
                s += t[d][1] * js[i][1]
                g.append(s)
            gp.append(g)
        return p, gp, abs(det)

    def _face(self, e, f):
        la, lb = FACE_V[f]
        va, vb = self.conn[e, la], self.conn[e, lb]
        pa, pb = self.xy[va], self.xy[vb]
        dx, dy = pb[0] - pa[0], pb[1] - pa[1]
        return la, lb, va, vb, pa, dx, dy, math.sqrt(dx * dx + dy * dy)

    # -- constraints / interpolation --
    def constraints_assemble(self, subproblems, bcs):
        self.constrained[:] = 0
        for e in range(self.ne):
            for f in range(3):
                nb = int(self.nbr[e, f])
                interior = nb >= 0
                la, lb, va, vb, pa, dx, dy, _ = self._face(e, f)
                xg = (pa[0] + 0.5 * dx, pa[1] + 0.5 * dy)
                for q, s in enumerate(subproblems):
                    if not self._applies(s, e):
                        continue
                    if interior and self._applies(s, nb):
                        continue
                    if eval_bctype(bcs[q], xg, not interior) != 0:
                        continue
                    ch = self.children[self.sps[s].child]
                    for v in (va, vb):
                        self.constrained[ch["offset"] + ch["v2d"][v]] = 1
        return int(self.constrained.sum())

    def get_constraints(self):
        return self.constrained.copy()

    def set_time(self, t):
        self.time = t

    def interpolate(self, subproblems, fns, x, time=0.0):
        for e in range(self.ne):
            for q, s in enumerate(subproblems):
                if not self._applies(s, e):
                    continue
                dofs = self._dofs(self.sps[s].child, e)
                for i in range(3):
                    x[dofs[i]] = eval_function(fns[q], self.xy[self.conn[e, i]], time)
        return x

    # -- operators --
    def gridoperator_create(self, subproblems, couplings=()):
        self.gos.append((list(subproblems), list(couplings)))
        return len(self.gos) - 1

    def _fires(self, q, e, f):
        nb = int(self.nbr[e, f])
        cp = self.cps[q]
        return nb >= 0 and self._applies(cp["local"], e) and self._applies(cp["remote"], nb)

    def matrix_create(self, gos):
        sps, cps = [], []
        for g in gos:
            sps += [s for s in self.gos[g][0] if s not in sps]
            cps += [q for q in self.gos[g][1] if q not in cps]
        rows = [set() for _ in range(self.ndof)]
        for e in range(self.ne):
            for s in sps:
                if self._applies(s, e):
                    d = self._dofs(self.sps[s].child, e)
                    for i in d:
                        rows[i].update(d)
            for f in range(3):
                for q in cps:
                    if self._fires(q, e, f):
                        nb = int(self.nbr[e, f])
                        ds = self._dofs(self.sps[self.cps[q]["local"]].child, e)
                        dn = self._dofs(self.sps[self.cps[q]["remote"]].child, nb)
                        for i in ds:
                            rows[i].update(dn)
                        for i in dn:
                            rows[i].update(ds)
        rowptr = np.zeros(self.ndof + 1, dtype=np.int64)
        cols = []
        for i, r in enumerate(rows):
            c = sorted(r)                                # [UPSTREAM] BCRS rows hold ascending column indices
            cols += c
            rowptr[i + 1] = len(cols)
        colidx = np.array(cols, dtype=np.int64)
        self.mats.append(dict(rowptr=rowptr, colidx=colidx, val=np.zeros(len(cols))))
        self.nnz = getattr(self, "nnz", {})
        self.nnz[len(self.mats) - 1] = len(cols)
        return len(self.mats) - 1

    def matrix_get_pattern(self, mat):
        return self.mats[mat]["rowptr"].copy(), self.mats[mat]["colidx"].copy()

    def matrix_get_values(self, mat):
        return self.mats[mat]["val"].copy()

    # ProportionalFlowCoupling::alpha_coupling (test/proportionalflowcoupling.hh:27-86), weighted local views
    def _propflow(self, F, k, xs, xn, weight, rs, rn):
        for phi1, phi2, factor in F:
            u_s = 0.0
            for i in range(3):
                u_s += xs[i] * phi1[i]
            u_n = 0.0
            for i in range(3):
                u_n += xn[i] * phi2[i]
            u_diff = k * (u_s - u_n)
            for i in range(3):
                rs[i] += weight * (u_diff * phi1[i] * factor)
            for i in range(3):
                rn[i] += weight * (-u_diff * phi2[i] * factor)

    # ContinuousValueContinuousFlowCoupling::alpha_coupling (test/proportionalflowcoupling.hh:113-233), theta = 1
    def _cvcf(self, F, G, alpha, x1, x2, weight, r1, r2):
        gp1, gp2, n1, h_F = G
        n2 = (-n1[0], -n1[1])
        for phi1, phi2, factor in F:
            gradu1, gradu2 = [0.0, 0.0], [0.0, 0.0]
            for i in range(3):
                for d in range(2):
                    gradu1[d] += x1[i] * gp1[i][d]
            for i in range(3):
                for d in range(2):
                    gradu2[d] += x2[i] * gp2[i][d]
            u1 = 0.0
            for i in range(3):
                u1 += x1[i] * phi1[i]
            u2 = 0.0
            for i in range(3):
                u2 += x2[i] * phi2[i]
            jump_u1, jump_u2 = u1 - u2, u2 - u1
            mean_gradu = [(gradu1[d] + gradu2[d]) * 0.5 for d in range(2)]
            for gp, n, phi, jump, r in ((gp1, n1, phi1, jump_u1, r1), (gp2, n2, phi2, jump_u2, r2)):
                for i in range(3):
                    mgn = 0.0
                    for d in range(2):
                        mgn += mean_gradu[d] * n[d]
                    gpn = 0.0
                    for d in range(2):
                        gpn += gp[i][d] * n[d]
                    value = 0.0
                    value -= mgn * phi[i]
                    value -= 0.5 * gpn * jump
                    value += alpha / h_F * jump * phi[i]
                    r[i] += weight * (factor * value)

    def _alpha_coupling(self, cp, e, f, xs, xn, weight, rs, rn):
        F = self._face_points(e, f, cp["order"])
        if cp["op"] == OP_CVCF:
            self._cvcf(F, self._face_gradients(e, f), cp["k"], xs, xn, weight, rs, rn)
        else:
            self._propflow(F, cp["k"], xs, xn, weight, rs, rn)

    def _face_gradients(self, e, f):
        nb = int(self.nbr[e, f])
        la, lb, va, vb, pa, dx, dy, ie = self._face(e, f)
        po = self.xy[self.conn[e, 2 - f]]                 # the inside cell's vertex opposite to the face
        nx, ny = dy / ie, -dx / ie
        side = nx * (po[0] - pa[0]) + ny * (po[1] - pa[1])
        if side > 0.0:
            nx, ny = -nx, -ny
        return self._tri(e)[1], self._tri(nb)[1], (nx, ny), ie

    def _face_points(self, e, f, order=4):
        nb = int(self.nbr[e, f])
        la, lb, va, vb, pa, dx, dy, ie = self._face(e, f)
        cn = list(self.conn[nb])
        ia, ib = cn.index(va), cn.index(vb)
        F = []
        for qx, qw in line_rule(order):
            x1 = REF[la][0] + qx * (REF[lb][0] - REF[la][0])
            y1 = REF[la][1] + qx * (REF[lb][1] - REF[la][1])
            x2 = REF[ia][0] + qx * (REF[ib][0] - REF[ia][0])
            y2 = REF[ia][1] + qx * (REF[ib][1] - REF[ia][1])
            F.append((p1_basis(x1, y1), p1_basis(x2, y2), qw * ie))
        return F

    def residual(self, go, x, r, weight=1.0):
        sps, cps = self.gos[go]
        for e in range(self.ne):
            for s in sps:
                if not self._applies(s, e):
                    continue
                sp = self.sps[s]
                dofs = self._dofs(sp.child, e)
                p, gp, ie = self._tri(e)
                xl = [x[d] for d in dofs]
                rl = [0.0, 0.0, 0.0]
                rule = tri_rule(sp.quad_order)
                for (qx, qy), w in rule:                                  # alpha_volume
                    factor = w * ie
                    phi = p1_basis(qx, qy)
                    if sp.op_id == OP_POISSON:
                        gradu = [0.0, 0.0]
                        for j in range(3):
                            for d in range(2):
                                gradu[d] += xl[j] * gp[j][d]
                        for i in range(3):
                            dot = 0.0
                            dot += gradu[0] * gp[i][0]
                            dot += gradu[1] * gp[i][1]
                            rl[i] += weight * (dot * factor)
                    else:
                        u = 0.0
                        for j in range(3):
                            u += xl[j] * phi[j]
                        for i in range(3):
                            rl[i] += weight * (u * phi[i] * factor)
                if sp.op_id == OP_POISSON and sp.P.f.kind != 0:           # lambda_volume
                    for (qx, qy), w in rule:
                        factor = w * ie
                        phi = p1_basis(qx, qy)
                        xg = [p[0][d] + (p[1][d] - p[0][d]) * qx + (p[2][d] - p[0][d]) * qy for d in range(2)]
                        y = eval_function(sp.P.f, xg, self.time)
                        for i in range(3):
                            rl[i] += weight * (-y * phi[i] * factor)
                if sp.op_id == OP_POISSON:                                # lambda_boundary: physical boundary faces and, with
                    for f in range(3):                                    # "no condition", interfaces (residualassemblerfunctors.hh:80-86)
                        nb = int(self.nbr[e, f])
                        if nb >= 0 and self._applies(s, nb):
                            continue
                        la, lb, va, vb, pa, dx, dy, fie = self._face(e, f)
                        for qx, qw in line_rule(sp.quad_order):
                            xg = (pa[0] + qx * dx, pa[1] + qx * dy)
                            if eval_bctype(sp.P.bc, xg, nb < 0) != 1:
                                continue
                            lx = REF[la][0] + qx * (REF[lb][0] - REF[la][0])
                            ly = REF[la][1] + qx * (REF[lb][1] - REF[la][1])
                            phi = p1_basis(lx, ly)
                            y = eval_function(sp.P.j, xg, self.time)
                            factor = qw * fie
                            for i in range(3):
                                rl[i] += weight * (y * phi[i] * factor)
                for i in range(3):
                    r[dofs[i]] += rl[i]
            for f in range(3):
                for q in cps:
                    if not self._fires(q, e, f):
                        continue
                    cp = self.cps[q]
                    nb = int(self.nbr[e, f])
                    ds = self._dofs(self.sps[cp["local"]].child, e)
                    dn = self._dofs(self.sps[cp["remote"]].child, nb)
                    rs, rn = [0.0] * 3, [0.0] * 3
                    self._alpha_coupling(cp, e, f, [x[d] for d in ds], [x[d] for d in dn], weight, rs, rn)
                    for i in range(3):
                        r[ds[i]] += rs[i]
                        r[dn[i]] += rn[i]
        r[self.constrained == 1] = 0.0                                   # constrain_residual (residualassemblerengine.hh:557)
        return r

    def jacobian(self, go, mat, x, weight=1.0, overwrite=False):
        M = self.mats[mat]
        rowptr, colidx, val = M["rowptr"], M["colidx"], M["val"]
        if overwrite:
            val[:] = 0.0

        def add(i, j, a):
            p0, p1 = rowptr[i], rowptr[i + 1]
            k = p0 + int(np.searchsorted(colidx[p0:p1], j))
            assert k < p1 and colidx[k] == j
            val[k] += a

        sps, cps = self.gos[go]
        for e in range(self.ne):
            for s in sps:
                if not self._applies(s, e):
                    continue
                sp = self.sps[s]
                dofs = self._dofs(sp.child, e)
                p, gp, ie = self._tri(e)
                a = [[0.0] * 3 for _ in range(3)]
                for (qx, qy), w in tri_rule(sp.quad_order):
                    factor = w * ie
                    phi = p1_basis(qx, qy)
                    for i in range(3):
                        for j in range(3):
                            if sp.op_id == OP_POISSON:
                                dot = 0.0
                                dot += gp[j][0] * gp[i][0]
                                dot += gp[j][1] * gp[i][1]
                                a[i][j] += weight * (dot * factor)
                            else:
                                a[i][j] += weight * (phi[j] * phi[i] * factor)
                for i in range(3):
                    for j in range(3):
                        add(dofs[i], dofs[j], a[i][j])
            for f in range(3):
                for q in cps:
                    if not self._fires(q, e, f):
                        continue
                    cp = self.cps[q]
                    nb = int(self.nbr[e, f])
                    g = self._dofs(self.sps[cp["local"]].child, e) + self._dofs(self.sps[cp["remote"]].child, nb)
                    analytic = cp["jac_mode"] == JAC_ANALYTIC
                    u = [0.0 if analytic else float(x[d]) for d in g]
                    down = [0.0] * 6
                    if not analytic:                                      # base line (couplingutilities.hh:94-97)
                        rs, rn = [0.0] * 3, [0.0] * 3
                        self._alpha_coupling(cp, e, f, u[:3], u[3:], 1.0, rs, rn)
                        down = rs + rn
                    for j in range(6):                                    # jiggle in self, then in the neighbour (:99-133)
                        keep = u[j]
                        delta = 1.0 if analytic else 1e-8 * (1.0 + abs(u[j]))
                        u[j] = 1.0 if analytic else u[j] + delta
                        rs, rn = [0.0] * 3, [0.0] * 3
                        self._alpha_coupling(cp, e, f, u[:3], u[3:], 1.0, rs, rn)
                        up = rs + rn
                        u[j] = keep
                        for i in range(6):
                            entry = up[i] if analytic else (up[i] - down[i]) / delta
                            add(g[i], g[j], weight * entry)
        for i in np.nonzero(self.constrained)[0]:                       # [UPSTREAM] handle_dirichlet_constraints: unit rows
            p0, p1 = rowptr[i], rowptr[i + 1]
            val[p0:p1] = 0.0
            k = p0 + int(np.searchsorted(colidx[p0:p1], i))
            if k < p1 and colidx[k] == i:
                val[k] = 1.0
