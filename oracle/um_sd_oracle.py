"""oracle/um_sd_oracle.py — TEST INFRASTRUCTURE ONLY (parity oracle of the Stokes-Darcy path, SURVEY §8 rows a5 + a12).

PARITY UNPINNED at the upstream boundary: the reference cannot be built here and its tests hold no expected values.
Restates, cell by cell in index order like GlobalAssembler::assemble (globalassembler.hh:43-298), what the reference's assembler does
for the driver test/stokesdarcy2.cc on a conforming triangle mesh:

  spaces        test/stokesdarcy2.cc:611-664: velocity = VectorGridFunctionSpace<SDGV, Pk(2), dim> and pressure = Pk(1) on the Stokes
                subdomain, composite (lexicographic) [v_0 | v_1 | p]; hydraulic head = OPBLocalFiniteElementMap<.., 3, 2, simplex> on the
                Darcy subdomain; MultiDomainGridFunctionSpace(lexicographic) -> child blocks [v_0 | v_1 | p | phi]
  numbering     multidomaingridfunctionspace.hh:308-369 with [UPSTREAM] (ii), (iii): per block geometry types in the order vertex < edge
                < cell, inside a type the rank of the entity among the subdomain's entities.  Edge indices of the host grid: order of first
                appearance in (cell, local face) order (UG's own edge numbering is an implementation detail of UG)
  local spaces  multidomainlocalfunctionspace.hh:348-358; the MULTI-CHILD subproblem view SubProblemLocalFunctionSpace
                (subproblemlocalfunctionspace.hh:144-261: the subproblem's children concatenated, remapped into the cell's local vector)
  pattern       patternassemblerengine.hh:142-187 (FullVolumePattern over ALL children of the subproblem, FullSkeletonPattern of the DG
                operator), couplingutilities.hh:35-47 (lfsv_s x lfsu_n and lfsv_n x lfsu_s over the whole subproblem spaces)
  operators     [UPSTREAM] TaylorHoodNavierStokes<P>(params, 4) with navier = false, full_tensor = true (test/stokesdarcy2.cc:233, :719-720),
                [UPSTREAM] ConvectionDiffusionDG<DarcyParams, OPB>(SIPG, weightsOn, alpha) (:733-739) with the tensor of DarcyParameters::A
                (:285-290), StokesDarcyCouplingOperator::alpha_coupling (test/stokesdarcycouplingoperator.hh:104-245) with the
                finite-difference Jacobian NumericalJacobianCoupling(epsilon) (couplingutilities.hh:75-135)
  skeleton      visited from the cell with the HIGHER index (globalassembler.hh:175-183: oneSidedDirection = ids > idn)
  constraints   constraints.hh:312-558; [UPSTREAM] StokesVelocityDirichletConstraints: VelocityDirichlet faces constrain the velocity
                coefficients of the face's sub-entities, the pressure never; the shipped StokesBoundaryType answers StressNeumann everywhere
                (test/stokesdarcy2.cc:139-140), so the shipped problem has no constrained DOF

[UPSTREAM] assumptions flagged: viscous form mu (grad u + grad u^T) : grad v, continuity rows -q div u, the source -f v, the stress
boundary term + (j n) . v with the scalar flux j (NavierStokesDefaultParameters), face rule of order stokes_q; the DG face terms as in
oracle/oracle.cpp (which restates them for cubes).  Exact Gauss rules: Gauss-Legendre on lines (order p -> ceil((p+1)/2) points), the
conical (collapsed) Gauss product rule with ceil((p+2)/2) points per direction on triangles — all integrands here are polynomials, so
any exact rule gives dune-geometry's integrals up to rounding.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.  Method names follow the ctypes binding of the
product (dune-multidomain_b200/capi.py) so that one problem-setup function drives both."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gen_opb_tables  # noqa: E402  (the offline table generator: the oracle evaluates the same rounded coefficients as the device)
from um_oracle import cond_applies

FACE_V = ((0, 1), (0, 2), (1, 2))
REF = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))
FE_P1, FE_P2, FE_OPB1, FE_OPB2, FE_OPB3 = 21, 22, 31, 32, 33
OP_TAYLORHOOD, OP_DARCY_DG, OP_SD_COUPLING = 4, 5, 104
JAC_FD, JAC_ANALYTIC = 0, 1


# ---- problem data (include/mdb200.h MDB_FN_*, MDB_BC_*) -------------------------------------------------------------------------
def eval_function(f, x):
    k, p = f.kind, f.p
    if k == 0:
        return 0.0
    if k == 1:
        return p[0]
    if k == 6:
        v, n2 = p[0], 0.0
        for d in range(2):
            v += p[1 + d] * x[d]
            n2 += x[d] * x[d]
        return v + p[4] * n2
    if k == 8:                                  # DarcyParameters::g  test/stokesdarcy2.cc:355-361
        scale = x[1] / p[2]
        return scale * p[0] + (1.0 - scale) * p[1]
    if k == 9:                                  # DarcyParameters::j  :364-371
        if x[1] < 1e-6:
            return p[0] * x[0] * (p[1] - x[0])
        return 0.0
    if k == 10:                                 # PressureDropFlux::evaluateGlobal  :191-203
        if x[0] < 1e-6:
            y = p[0] if x[1] > p[3] else -p[1]
        elif x[0] > p[2] - 1e-6:
            y = -p[1] if x[1] > p[3] else p[0]
        else:
            y = 0.0
        return y * (p[4] - x[1])
    raise ValueError("function kind %d not restated here" % k)


def eval_bctype(b, x, is_boundary):
    """0 Dirichlet (VelocityDirichlet), 1 Neumann / Outflow (StressNeumann), 2 none (DoNothing)."""
    if not is_boundary:
        return 2
    if b.kind == 0:
        return 0
    if b.kind == 1:
        return 1
    if b.kind == 3:                             # DarcyParameters::bctype  test/stokesdarcy2.cc:338-353 (Outflow with b = 0 acts like Neumann with o = j)
        if x[1] < 1e-6 and x[0] < b.p[0]:
            return 1
        if x[0] > b.p[0] - 1e-6 and x[1] < b.p[1]:
            return 0
        return 1
    raise ValueError("bc kind %d not restated here" % b.kind)


# ---- rules ---------------------------------------------------------------------------------------------------------------------------
def gauss01(m):
    """Gauss-Legendre on [0, 1] in closed form (correctly rounded sqrt / division): the finite-difference coupling Jacobian with a small
    epsilon amplifies a 1-ulp difference in a node to 1e-8 relative in the entry, so the rule constants are spelled out, not iterated."""
    sq = math.sqrt
    if m == 1:
        return [(0.5, 1.0)]
    if m == 2:
        a = 0.5 / sq(3.0)
        return [(0.5 - a, 0.5), (0.5 + a, 0.5)]
    if m == 3:
        a = 0.5 * sq(3.0 / 5.0)
        return [(0.5 - a, 5.0 / 18.0), (0.5, 8.0 / 18.0), (0.5 + a, 5.0 / 18.0)]
    if m == 4:
        s = sq(6.0 / 5.0)
        a = 0.5 * sq(3.0 / 7.0 - 2.0 / 7.0 * s)
        b = 0.5 * sq(3.0 / 7.0 + 2.0 / 7.0 * s)
        wa = (18.0 + sq(30.0)) / 72.0
        wb = (18.0 - sq(30.0)) / 72.0
        return [(0.5 - b, wb), (0.5 - a, wa), (0.5 + a, wa), (0.5 + b, wb)]
    if m == 5:
        s = 2.0 * sq(10.0 / 7.0)
        a = 0.5 * (1.0 / 3.0) * sq(5.0 - s)
        b = 0.5 * (1.0 / 3.0) * sq(5.0 + s)
        wa = (322.0 + 13.0 * sq(70.0)) / 1800.0
        wb = (322.0 - 13.0 * sq(70.0)) / 1800.0
        return [(0.5 - b, wb), (0.5 - a, wa), (0.5, 128.0 / 450.0), (0.5 + a, wa), (0.5 + b, wb)]
    raise ValueError("Gauss rules up to 5 points")


def line_rule(order):
    return gauss01((order + 2) // 2)


def tri_rule(order):
    g = gauss01((order + 3) // 2)
    return [((s, t * (1.0 - s)), ws * wt * (1.0 - s)) for s, ws in g for t, wt in g]


# ---- local bases -----------------------------------------------------------------------------------------------------------------------
def pk_eval(k, x, y):
    """Pk2D Lagrange basis (dune-localfunctions Pk2DLocalBasis [UPSTREAM]): values and reference gradients, lattice order, i fastest."""
    pos = [a / float(k) for a in range(k + 1)]
    val, grad = [], []
    for j in range(k + 1):
        for i in range(k + 1 - j):
            fac = [((x - pos[a]) / (pos[i] - pos[a]), 1.0 / (pos[i] - pos[a]), 0.0) for a in range(i)]
            fac += [((y - pos[b]) / (pos[j] - pos[b]), 0.0, 1.0 / (pos[j] - pos[b])) for b in range(j)]
            fac += [((pos[c] - x - y) / (pos[c] - pos[i] - pos[j]), -1.0 / (pos[c] - pos[i] - pos[j]), -1.0 / (pos[c] - pos[i] - pos[j]))
                    for c in range(i + j + 1, k + 1)]
            v = 1.0
            for f in fac:
                v *= f[0]
            gx = gy = 0.0
            for m in range(len(fac)):
                px = py = 1.0
                for l in range(len(fac)):
                    if l == m:
                        px *= fac[l][1]
                        py *= fac[l][2]
                    else:
                        px *= fac[l][0]
                        py *= fac[l][0]
                gx += px
                gy += py
            val.append(v)
            grad.append((gx, gy))
    return np.array(val), np.array(grad)


_OPB = {}


def opb_eval(k, x, y):
    """Orthonormal basis of degree k (gen_opb_tables): phi_i = sum_{j<=i} L_ij x^a y^b, summed in j order; powers by repeated products."""
    if k not in _OPB:
        _OPB[k] = gen_opb_tables.opb_matrix(k)
    L, mon = _OPB[k]
    n = len(mon)
    px, py = [1.0], [1.0]
    for _ in range(k):
        px.append(px[-1] * x)
        py.append(py[-1] * y)
    m = [px[a] * py[b] for a, b in mon]
    mx = [(a * px[a - 1] * py[b]) if a > 0 else 0.0 for a, b in mon]
    my = [(b * px[a] * py[b - 1]) if b > 0 else 0.0 for a, b in mon]
    val, grad = np.zeros(n), np.zeros((n, 2))
    for i in range(n):
        v = gx = gy = 0.0
        for j in range(i + 1):
            v += L[i][j] * m[j]
            gx += L[i][j] * mx[j]
            gy += L[i][j] * my[j]
        val[i], grad[i, 0], grad[i, 1] = v, gx, gy
    return val, grad


def fe_eval(fe, x, y):
    if fe == FE_P1:
        return pk_eval(1, x, y)
    if fe == FE_P2:
        return pk_eval(2, x, y)
    return opb_eval(fe - 30, x, y)


def fe_size(fe):
    return {FE_P1: 3, FE_P2: 6, FE_OPB1: 3, FE_OPB2: 6, FE_OPB3: 10}[fe]


def fe_order(fe):
    return {FE_P1: 1, FE_P2: 2, FE_OPB1: 1, FE_OPB2: 2, FE_OPB3: 3}[fe]


class _Sp:
    pass


class SDOracle:
    """CPU restatement; same call shapes as dune-multidomain_b200.capi.Mdb for the generic unstructured path."""

    def __init__(self):
        self.children, self.sps, self.cps, self.gos, self.mats = [], [], [], [], []
        self.groups = None
        self.nnz = {}

    # -- mesh --
    def mesh_unstructured(self, xy, conn):
        self.xy = np.asarray(xy, dtype=np.float64)
        self.conn = np.asarray(conn, dtype=np.int64)
        self.nv, self.ne = self.xy.shape[0], self.conn.shape[0]
        self.ncells = self.ne
        edge, first = {}, {}
        self.nbr = -np.ones((self.ne, 3), dtype=np.int64)
        self.cell_edge = np.zeros((self.ne, 3), dtype=np.int64)
        for e in range(self.ne):
            for f in range(3):
                a, b = self.conn[e, FACE_V[f][0]], self.conn[e, FACE_V[f][1]]
                key = (min(a, b), max(a, b))
                if key in edge:
                    e2, f2 = edge[key]
                    self.nbr[e, f], self.nbr[e2, f2] = e2, e
                else:
                    edge[key] = (e, f)
                    first[key] = len(first)                       # host edge index: order of first appearance
                self.cell_edge[e, f] = first[key]
        self.nedges = len(first)
        self.groups = np.zeros(self.ne, dtype=np.int64)

    def subdomains(self, bits):
        self.sd = np.asarray(bits, dtype=np.uint8).astype(np.int64)

    def cell_groups(self, groups):
        self.groups = np.asarray(groups, dtype=np.int64)

    def add_space(self, fe_kind, subdomain):
        self.children.append(dict(fe=fe_kind, subdomain=subdomain))
        return len(self.children) - 1

    def add_subproblem(self, op_id, params, cond_kind, cond_mask, children):
        s = _Sp()
        s.op_id, s.P, s.kind, s.mask, s.children = op_id, params, cond_kind, cond_mask, [int(c) for c in children]
        self.sps.append(s)
        return len(self.sps) - 1

    def add_coupling(self, op_id, params, local_sp, remote_sp, jac_mode=JAC_FD):
        assert op_id == OP_SD_COUPLING
        self.cps.append(dict(op=op_id, P=params, local=local_sp, remote=remote_sp, jac_mode=jac_mode))
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
            cells = [e for e in range(self.ne) if self._present(k, e)]
            vflag, eflag = np.zeros(self.nv, dtype=bool), np.zeros(self.nedges, dtype=bool)
            for e in cells:
                vflag[self.conn[e]] = True
                eflag[self.cell_edge[e]] = True
            v2d = -np.ones(self.nv, dtype=np.int64); v2d[vflag] = np.arange(int(vflag.sum()))
            e2d = -np.ones(self.nedges, dtype=np.int64); e2d[eflag] = np.arange(int(eflag.sum()))
            c2d = -np.ones(self.ne, dtype=np.int64); c2d[cells] = np.arange(len(cells))
            fe = ch["fe"]
            if fe == FE_P1:
                size = int(vflag.sum())
            elif fe == FE_P2:
                size = int(vflag.sum()) + int(eflag.sum())
            else:
                size = len(cells) * fe_size(fe)
            ch.update(v2d=v2d, e2d=e2d, c2d=c2d, nvs=int(vflag.sum()), size=size, offset=off)
            off += size
        self.ndof = off
        self.constrained = np.zeros(self.ndof, dtype=np.uint8)
        return self.ndof

    def _child_dofs(self, k, e):
        ch = self.children[k]
        fe, off = ch["fe"], ch["offset"]
        c = self.conn[e]
        if fe == FE_P1:
            return [off + int(ch["v2d"][v]) for v in c]
        if fe == FE_P2:
            # Pk2D lattice order (0,0),(1/2,0),(1,0),(0,1/2),(1/2,1/2),(0,1): vertex 0, edge {0,1}, vertex 1, edge {0,2}, edge {1,2}, vertex 2
            ed = [ch["nvs"] + int(ch["e2d"][g]) for g in self.cell_edge[e]]
            vv = [int(ch["v2d"][v]) for v in c]
            return [off + i for i in (vv[0], ed[0], vv[1], ed[1], ed[2], vv[2])]
        n = fe_size(fe)
        return [off + int(ch["c2d"][e]) * n + i for i in range(n)]

    def _sp_dofs(self, s, e):
        out = []
        for k in self.sps[s].children:
            out += self._child_dofs(k, e)
        return out

    def get_dofmap(self):
        ptr, dofs = [0], []
        for e in range(self.ne):
            for k in range(len(self.children)):
                if self._present(k, e):
                    dofs += self._child_dofs(k, e)
            ptr.append(len(dofs))
        return np.array(ptr, dtype=np.int64), np.array(dofs, dtype=np.int64)

    def get_child_offsets(self):
        return np.array([ch["offset"] for ch in self.children] + [self.ndof], dtype=np.int64)

    # -- geometry --
    def _geo(self, e):
        p = self.xy[self.conn[e]]
        ax, ay = p[1][0] - p[0][0], p[1][1] - p[0][1]
        bx, by = p[2][0] - p[0][0], p[2][1] - p[0][1]
        det = ax * by - ay * bx
        inv = 1.0 / det
        t = np.array([[by * inv, -(ay * inv)], [-(bx * inv), ax * inv]])          # J^{-T}
        return p, t, abs(det)

    def _face(self, e, f):
        """(la, lb, pa, edge vector, length, unit outer normal of cell e)"""
        la, lb = FACE_V[f]
        p = self.xy[self.conn[e]]
        pa, pb, po = p[la], p[lb], p[2 - f]
        dx, dy = pb[0] - pa[0], pb[1] - pa[1]
        ie = math.sqrt(dx * dx + dy * dy)
        nx, ny = dy / ie, -dx / ie
        if nx * (po[0] - pa[0]) + ny * (po[1] - pa[1]) > 0.0:
            nx, ny = -nx, -ny
        return la, lb, pa, (dx, dy), ie, (nx, ny)

    def _opposite(self, e, f, nb):
        """local vertex indices in the neighbour of the face's two vertices (geometryInOutside)"""
        la, lb = FACE_V[f]
        cn = list(self.conn[nb])
        return cn.index(self.conn[e, la]), cn.index(self.conn[e, lb])

    @staticmethod
    def _ref_point(a, b, t):
        return REF[a][0] + t * (REF[b][0] - REF[a][0]), REF[a][1] + t * (REF[b][1] - REF[a][1])

    # -- constraints --
    def constraints_assemble(self, subproblems, bcs):
        self.constrained[:] = 0
        for e in range(self.ne):
            for f in range(3):
                nb = int(self.nbr[e, f])
                la, lb, pa, (dx, dy), ie, n = self._face(e, f)
                xg = (pa[0] + 0.5 * dx, pa[1] + 0.5 * dy)
                for q, s in enumerate(subproblems):
                    sp = self.sps[s]
                    if not self._applies(s, e) or (nb >= 0 and self._applies(s, nb)):
                        continue
                    if eval_bctype(bcs[q], xg, nb < 0) != 0:
                        continue
                    # sub-entities of the face in the P2 local numbering: vertices la, lb and the edge f
                    p2_face = {0: (0, 1, 2), 1: (0, 3, 5), 2: (2, 4, 5)}[f]
                    kids = sp.children[:2] if sp.op_id == OP_TAYLORHOOD else sp.children      # the pressure is never constrained
                    for k in kids:
                        fe = self.children[k]["fe"]
                        d = self._child_dofs(k, e)
                        if fe == FE_P2:
                            for i in p2_face:
                                self.constrained[d[i]] = 1
                        elif fe == FE_P1:
                            self.constrained[d[la]] = 1
                            self.constrained[d[lb]] = 1
        return int(self.constrained.sum())

    def get_constraints(self):
        return self.constrained.copy()

    def set_time(self, t):
        pass

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
                if not self._applies(s, e):
                    continue
                d = self._sp_dofs(s, e)
                for i in d:
                    rows[i].update(d)
                if self.sps[s].op_id == OP_DARCY_DG:                       # doPatternSkeleton: FullSkeletonPattern
                    for f in range(3):
                        nb = int(self.nbr[e, f])
                        if nb >= 0 and self._applies(s, nb):
                            dn = self._sp_dofs(s, nb)
                            for i in d:
                                rows[i].update(dn)
            for f in range(3):
                for q in cps:
                    if self._fires(q, e, f):
                        nb = int(self.nbr[e, f])
                        ds, dn = self._sp_dofs(self.cps[q]["local"], e), self._sp_dofs(self.cps[q]["remote"], nb)
                        for i in ds:
                            rows[i].update(dn)
                        for i in dn:
                            rows[i].update(ds)
        rowptr = np.zeros(self.ndof + 1, dtype=np.int64)
        cols = []
        for i, r in enumerate(rows):
            cols += sorted(r)
            rowptr[i + 1] = len(cols)
        self.mats.append(dict(rowptr=rowptr, colidx=np.array(cols, dtype=np.int64), val=np.zeros(len(cols))))
        self.nnz[len(self.mats) - 1] = len(cols)
        return len(self.mats) - 1

    def matrix_get_pattern(self, mat):
        return self.mats[mat]["rowptr"].copy(), self.mats[mat]["colidx"].copy()

    def matrix_get_values(self, mat):
        return self.mats[mat]["val"].copy()

    def matrix_zero(self, mat):
        self.mats[mat]["val"][:] = 0.0

    # ---- [UPSTREAM] TaylorHoodNavierStokes (Stokes, full tensor): alpha_volume + lambda_volume + lambda_boundary on cell e ----
    def _th_local(self, sp, e, xl):
        P = sp.P
        p, t, ie = self._geo(e)
        xv = [xl[0:6], xl[6:12]]
        xp = xl[12:15]
        r = np.zeros(15)
        for (lx, ly), w in tri_rule(P.quad_order):
            phi, js = pk_eval(2, lx, ly)
            psi, _ = pk_eval(1, lx, ly)
            grad = js @ t.T                                              # J^{-T} applied to every reference gradient
            jacu = [xv[0] @ grad, xv[1] @ grad]
            pres = float(xp @ psi)
            divu = jacu[0][0] + jacu[1][1]
            factor = w * ie
            for d in range(2):
                r[6 * d:6 * d + 6] += (P.mu * (grad @ jacu[d]) - pres * grad[:, d]) * factor
                if P.full_tensor:
                    for dd in range(2):
                        r[6 * d:6 * d + 6] += P.mu * jacu[dd][d] * grad[:, dd] * factor
                r[6 * d:6 * d + 6] += -P.f[d] * phi * factor              # lambda_volume
            r[12:15] += -divu * psi * factor
        return r

    def _th_boundary(self, sp, s, e):
        """lambda_boundary: stress boundary faces (physical boundary; on a subdomain interface the type is DoNothing)"""
        P = sp.P
        r = np.zeros(15)
        for f in range(3):
            nb = int(self.nbr[e, f])
            if nb >= 0 and self._applies(s, nb):
                continue
            la, lb, pa, (dx, dy), fie, n = self._face(e, f)
            for tq, w in line_rule(P.quad_order):
                xg = (pa[0] + tq * dx, pa[1] + tq * dy)
                if eval_bctype(P.bc, xg, nb < 0) != 1:
                    continue
                phi, _ = pk_eval(2, *self._ref_point(la, lb, tq))
                jv = eval_function(P.j, xg)
                factor = w * fie
                for d in range(2):
                    r[6 * d:6 * d + 6] += (jv * n[d]) * phi * factor
        return r

    # ---- [UPSTREAM] ConvectionDiffusionDG with DarcyParameters: volume + skeleton + boundary ----
    def _A(self, sp, e):
        k = sp.P.kabs if int(self.groups[e]) == sp.P.high_group else sp.P.kabs_low
        return np.array([[k, 0.0], [0.0, k]])

    def _dg_volume(self, sp, fe, e, xs):
        p, t, ie = self._geo(e)
        A = self._A(sp, e)
        order = sp.P.intorderadd + 2 * fe_order(fe)
        r = np.zeros(fe_size(fe))
        for (lx, ly), w in tri_rule(order):
            phi, js = fe_eval(fe, lx, ly)
            grad = js @ t.T
            Agradu = A @ (xs @ grad)
            r += (grad @ Agradu) * (w * ie)                              # f = 0, c = 0, b = 0 (test/stokesdarcy2.cc:277-335)
        return r

    def _dg_face(self, sp, s, fe, e, f, xs, xn):
        """Face f of cell e seen from e (inside).  Returns (r_s, r_n); r_n is None on a boundary face."""
        P = sp.P
        k = fe_order(fe)
        order = P.intorderadd + 2 * k
        theta = 1.0 if P.method == 0 else -1.0
        nb = int(self.nbr[e, f])
        boundary = nb < 0
        la, lb, pa, (dx, dy), fie, n = self._face(e, f)
        nF = np.array(n)
        ps, ts, ies = self._geo(e)
        A_s = self._A(sp, e)
        An_s = A_s @ nF
        delta_s = float(An_s @ nF)
        ns = fe_size(fe)
        r_s, r_n = np.zeros(ns), None
        if boundary:
            bc = eval_bctype(P.bc, (pa[0] + 0.5 * dx, pa[1] + 0.5 * dy), True)           # at the face centre
            omega_s, omega_n, harmonic, vol = 1.0, 0.0, delta_s, 0.5 * ies
        else:
            pn, tn, ien = self._geo(nb)
            An_n = self._A(sp, nb) @ nF
            delta_n = float(An_n @ nF)
            if P.weights:
                omega_s, omega_n = delta_n / (delta_s + delta_n + 1e-20), delta_s / (delta_s + delta_n + 1e-20)
                harmonic = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20)
            else:
                omega_s = omega_n = 0.5
                harmonic = 1.0
            vol = min(0.5 * ies, 0.5 * ien)
            ia, ib = self._opposite(e, f, nb)
            r_n = np.zeros(ns)
        if boundary and not P.weights:
            harmonic = 1.0
        h_F = vol / fie
        penalty = (P.alpha / h_F) * harmonic * k * (k + 2 - 1)
        for tq, w in line_rule(order):
            factor = w * fie
            phi_s, js = fe_eval(fe, *self._ref_point(la, lb, tq))
            grad_s = js @ ts.T
            xg = (pa[0] + tq * dx, pa[1] + tq * dy)
            u_s, gradu_s = float(xs @ phi_s), xs @ grad_s
            if boundary:
                if bc == 1:
                    r_s += eval_function(P.j, xg) * phi_s * factor
                elif bc == 0:
                    gval = eval_function(P.g, xg)
                    r_s += -float(An_s @ gradu_s) * factor * phi_s
                    r_s += (u_s - gval) * factor * theta * (grad_s @ An_s)
                    r_s += penalty * (u_s - gval) * factor * phi_s
                continue
            phi_n, jn = fe_eval(fe, *self._ref_point(ia, ib, tq))
            grad_n = jn @ tn.T
            u_n, gradu_n = float(xn @ phi_n), xn @ grad_n
            term2 = -(omega_s * float(An_s @ gradu_s) + omega_n * float(An_n @ gradu_n)) * factor
            term3 = (u_s - u_n) * factor
            term4 = penalty * (u_s - u_n) * factor
            r_s += term2 * phi_s + term3 * theta * omega_s * (grad_s @ An_s) + term4 * phi_s
            r_n += -term2 * phi_n + term3 * theta * omega_n * (grad_n @ An_n) - term4 * phi_n
        return r_s, r_n

    # ---- StokesDarcyCouplingOperator::alpha_coupling (test/stokesdarcycouplingoperator.hh:104-245), statement by statement ----
    def _sd_coupling(self, cp, e, f, xs, xn):
        P = cp["P"]
        nb = int(self.nbr[e, f])
        spL, spR = self.sps[cp["local"]], self.sps[cp["remote"]]
        fe_v, fe_d = self.children[spL.children[0]]["fe"], self.children[spR.children[0]]["fe"]
        vsize, dsize = fe_size(fe_v), fe_size(fe_d)
        la, lb, pa, (dx, dy), fie, n = self._face(e, f)
        ia, ib = self._opposite(e, f, nb)
        pn, tn, ien = self._geo(nb)
        qorder = 2 * max(fe_order(fe_v), fe_order(fe_d))                                   # :159-160
        kabs = ((P.kabs, 0.0), (0.0, P.kabs))
        tracePi = 0.0
        for i in range(2):
            tracePi += kabs[i][i]
        tracePi *= P.viscosity / P.gravity / P.density                                     # :172-175
        nls = len(xs)
        r_s, r_n = np.zeros(nls), np.zeros(dsize)
        for tq, w in line_rule(qorder):
            pos = (pa[0] + tq * dx, pa[1] + tq * dy)
            factor = w * fie
            v, _ = fe_eval(fe_v, *self._ref_point(la, lb, tq))
            psi, js = fe_eval(fe_d, *self._ref_point(ia, ib, tq))
            gradpsi = np.zeros((dsize, 2))                                                 # value-initialised FieldVectors
            for i in range(vsize if P.quirk else dsize):                                   # :194-195 loops i < vsize
                if i < dsize:                                                              # jac.mv(js[i][0], gradpsi[i]), scalar order
                    gradpsi[i, 0] = tn[0][0] * js[i][0] + tn[0][1] * js[i][1]
                    gradpsi[i, 1] = tn[1][0] * js[i][0] + tn[1][1] * js[i][1]
            phi, gradphi = 0.0, np.zeros(2)
            for i in range(dsize):
                phi += xn[i] * psi[i]
                gradphi += xn[i] * gradpsi[i]
            u = np.zeros(2)
            for d in range(2):
                for i in range(vsize):
                    u[d] += xs[vsize * d + i] * v[i]
            un = u[0] * n[0] + u[1] * n[1]
            for i in range(dsize):                                                         # :219
                r_n[i] += -P.gamma * P.porosity * un * psi[i] * factor
            tang = np.array([kabs[0][0] * gradphi[0] + kabs[0][1] * gradphi[1], kabs[1][0] * gradphi[0] + kabs[1][1] * gradphi[1]])
            tang /= P.porosity
            tang += u
            tn_ = tang[0] * n[0] + tang[1] * n[1]
            tang -= np.array([n[0] * tn_, n[1] * tn_])                                     # project into the tangential plane
            for d in range(2):
                for i in range(vsize):
                    r_s[vsize * d + i] += -P.density * P.gravity * (phi - pos[1]) * v[i] * n[d] * factor          # :235
                for i in range(vsize):
                    r_s[vsize * d + i] += P.alpha * math.sqrt(2.0) / math.sqrt(tracePi) * tang[d] * v[i] * factor   # :240 Beavers-Joseph
        return r_s, r_n

    # ---- assembly ----
    def _volume_terms(self, s, e, x):
        """alpha_volume + lambda_* of subproblem s on cell e at the coefficients x (global vector): local residual over _sp_dofs."""
        sp = self.sps[s]
        d = self._sp_dofs(s, e)
        xl = x[d]
        if sp.op_id == OP_TAYLORHOOD:
            return self._th_local(sp, e, xl) + self._th_boundary(sp, s, e)
        fe = self.children[sp.children[0]]["fe"]
        return self._dg_volume(sp, fe, e, xl)

    def residual(self, go, x, r, weight=1.0):
        sps, cps = self.gos[go]
        x = np.asarray(x, dtype=np.float64)
        for e in range(self.ne):
            for s in sps:
                if not self._applies(s, e):
                    continue
                sp = self.sps[s]
                d = self._sp_dofs(s, e)
                r[d] += weight * self._volume_terms(s, e, x)
                if sp.op_id == OP_DARCY_DG:
                    fe = self.children[sp.children[0]]["fe"]
                    for f in range(3):
                        nb = int(self.nbr[e, f])
                        if nb >= 0 and self._applies(s, nb):
                            if e < nb:
                                continue                                  # skeleton term from the higher-index cell
                            dn = self._sp_dofs(s, nb)
                            rs, rn = self._dg_face(sp, s, fe, e, f, x[d], x[dn])
                            r[d] += weight * rs
                            r[dn] += weight * rn
                        elif nb < 0:
                            rs, _ = self._dg_face(sp, s, fe, e, f, x[d], None)
                            r[d] += weight * rs
                        # else: interface seen as a boundary with BCType None (test/stokesdarcy2.cc:341-342): no term
            for f in range(3):
                for q in cps:
                    if not self._fires(q, e, f):
                        continue
                    cp = self.cps[q]
                    nb = int(self.nbr[e, f])
                    ds, dn = self._sp_dofs(cp["local"], e), self._sp_dofs(cp["remote"], nb)
                    rs, rn = self._sd_coupling(cp, e, f, x[ds], x[dn])
                    r[ds] += weight * rs
                    r[dn] += weight * rn
        r[self.constrained == 1] = 0.0
        return r

    def jacobian(self, go, mat, x, weight=1.0, overwrite=False):
        M = self.mats[mat]
        rowptr, colidx, val = M["rowptr"], M["colidx"], M["val"]
        if overwrite:
            val[:] = 0.0
        x = np.asarray(x, dtype=np.float64)

        def add_block(rows, cols, B):
            for a, i in enumerate(rows):
                p0, p1 = rowptr[i], rowptr[i + 1]
                pos = p0 + np.searchsorted(colidx[p0:p1], cols)
                assert np.array_equal(colidx[pos], np.asarray(cols))
                np.add.at(val, pos, B[a])

        sps, cps = self.gos[go]
        for e in range(self.ne):
            for s in sps:
                if not self._applies(s, e):
                    continue
                sp = self.sps[s]
                d = self._sp_dofs(s, e)
                n = len(d)
                # the volume operators are affine: the analytic jacobian_volume = columns alpha(e_j) - alpha(0)
                z = np.zeros(n)
                if sp.op_id == OP_TAYLORHOOD:
                    base = self._th_local(sp, e, z)
                    K = np.stack([self._th_local(sp, e, np.eye(n)[j]) - base for j in range(n)], axis=1)
                else:
                    fe = self.children[sp.children[0]]["fe"]
                    K = np.stack([self._dg_volume(sp, fe, e, np.eye(n)[j]) for j in range(n)], axis=1)
                add_block(d, d, weight * K)
                if sp.op_id == OP_DARCY_DG:
                    for f in range(3):
                        nb = int(self.nbr[e, f])
                        if nb >= 0 and self._applies(s, nb):
                            if e < nb:
                                continue
                            dn = self._sp_dofs(s, nb)
                            b_s, b_n = self._dg_face(sp, s, fe, e, f, z, z)
                            ss, ns_, sn, nn = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
                            for j in range(n):
                                rs, rn = self._dg_face(sp, s, fe, e, f, np.eye(n)[j], z)
                                ss[:, j], ns_[:, j] = rs - b_s, rn - b_n
                                rs, rn = self._dg_face(sp, s, fe, e, f, z, np.eye(n)[j])
                                sn[:, j], nn[:, j] = rs - b_s, rn - b_n
                            add_block(d, d, weight * ss); add_block(d, dn, weight * sn)
                            add_block(dn, d, weight * ns_); add_block(dn, dn, weight * nn)
                        elif nb < 0:
                            b_s, _ = self._dg_face(sp, s, fe, e, f, z, None)
                            ss = np.stack([self._dg_face(sp, s, fe, e, f, np.eye(n)[j], None)[0] - b_s for j in range(n)], axis=1)
                            add_block(d, d, weight * ss)
            for f in range(3):
                for q in cps:
                    if not self._fires(q, e, f):
                        continue
                    cp = self.cps[q]
                    nb = int(self.nbr[e, f])
                    ds, dn = self._sp_dofs(cp["local"], e), self._sp_dofs(cp["remote"], nb)
                    n_s, n_n = len(ds), len(dn)
                    analytic = cp["jac_mode"] == JAC_ANALYTIC
                    eps = cp["P"].epsilon
                    u_s = np.zeros(n_s) if analytic else x[ds].copy()
                    u_n = np.zeros(n_n) if analytic else x[dn].copy()
                    down_s, down_n = self._sd_coupling(cp, e, f, u_s, u_n)                # base line (couplingutilities.hh:94-97)
                    ss, ns_, sn, nn = np.zeros((n_s, n_s)), np.zeros((n_n, n_s)), np.zeros((n_s, n_n)), np.zeros((n_n, n_n))
                    for j in range(n_s):                                                   # jiggle in self (:99-113)
                        keep = u_s[j]
                        delta = 1.0 if analytic else eps * (1.0 + abs(u_s[j]))
                        u_s[j] += delta
                        up_s, up_n = self._sd_coupling(cp, e, f, u_s, u_n)
                        ss[:, j], ns_[:, j] = (up_s - down_s) / delta, (up_n - down_n) / delta
                        u_s[j] = keep
                    for j in range(n_n):                                                   # jiggle in neighbour (:115-131)
                        keep = u_n[j]
                        delta = 1.0 if analytic else eps * (1.0 + abs(u_n[j]))
                        u_n[j] += delta
                        up_s, up_n = self._sd_coupling(cp, e, f, u_s, u_n)
                        sn[:, j], nn[:, j] = (up_s - down_s) / delta, (up_n - down_n) / delta
                        u_n[j] = keep
                    add_block(ds, ds, weight * ss); add_block(ds, dn, weight * sn)
                    add_block(dn, ds, weight * ns_); add_block(dn, dn, weight * nn)
        for i in np.nonzero(self.constrained)[0]:                                          # [UPSTREAM] handle_dirichlet_constraints
            p0, p1 = rowptr[i], rowptr[i + 1]
            val[p0:p1] = 0.0
            k = p0 + int(np.searchsorted(colidx[p0:p1], i))
            if k < p1 and colidx[k] == i:
                val[k] = 1.0
