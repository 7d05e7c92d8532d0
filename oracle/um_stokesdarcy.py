"""oracle/um_stokesdarcy.py — TEST INFRASTRUCTURE ONLY.  Groundwork for SURVEY §8 row a12 (no device path yet).

Restates, for one interface edge of a triangle mesh, the reference's Stokes-Darcy coupling operator
StokesDarcyCouplingOperator::alpha_coupling (test/stokesdarcycouplingoperator.hh:104-245) and the local bases it is evaluated
with: Taylor-Hood velocity P2 per component on the Stokes (inside) cell, DG-P3 hydraulic head on the Darcy (outside) cell
(test/stokesdarcy2.cc:613-616, :646-700).  PARITY UNPINNED at the upstream boundary like every oracle in this directory; pinned by
tests/test_oracle_stokesdarcy.py (basis identities, closed-form row sums, hydrostatic equilibrium).  The second half restates the
Darcy subproblem's operator, [UPSTREAM] ConvectionDiffusionDG(SIPG, weightsOn) on DG-Pk triangles with a tensor (class DGTriangles).

[UPSTREAM] dune-localfunctions Pk2DLocalBasis<D,R,k>: Lagrange basis on the lattice (i/k, j/k), i + j <= k, numbered with i fastest:
    phi_n(x, y) = prod_{a<i} (x - a/k)/(i/k - a/k) * prod_{b<j} (y - b/k)/(j/k - b/k) * prod_{c=i+j+1..k} (c/k - x - y)/(c/k - i/k - j/k)

Reference quirk restated as written (SURVEY §8a "quirks"): the loop that transforms the Darcy gradients to the real element runs over
`vsize` (6, the number of P2 velocity functions) instead of `dsize` (10), stokesdarcycouplingoperator.hh:194-195, so the gradients
of the last dsize - vsize head functions stay at their initial value; `quirk=False` gives the evidently intended operator.  The
quirk matters: with it the hydrostatic state (u = 0, phi = z) is NOT an equilibrium of the coupling terms (tests pin both).
"""
import math

import numpy as np

FACE_V = ((0, 1), (0, 2), (1, 2))
REF = ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0))


def pk_size(k):
    return (k + 1) * (k + 2) // 2


def pk_basis(k, x, y):
    """Values of the Pk2D basis at (x, y)."""
    pos = [a / float(k) for a in range(k + 1)]
    out = []
    for j in range(k + 1):
        for i in range(k + 1 - j):
            v = 1.0
            for a in range(i):
                v *= (x - pos[a]) / (pos[i] - pos[a])
            for b in range(j):
                v *= (y - pos[b]) / (pos[j] - pos[b])
            for c in range(i + j + 1, k + 1):
                v *= (pos[c] - x - y) / (pos[c] - pos[i] - pos[j])
            out.append(v)
    return out


def pk_gradients(k, x, y):
    """Reference gradients (d/dx, d/dy) of the Pk2D basis at (x, y): product rule over the three groups of factors."""
    pos = [a / float(k) for a in range(k + 1)]
    out = []
    for j in range(k + 1):
        for i in range(k + 1 - j):
            fac = [((x - pos[a]) / (pos[i] - pos[a]), 1.0 / (pos[i] - pos[a]), 0.0) for a in range(i)]
            fac += [((y - pos[b]) / (pos[j] - pos[b]), 0.0, 1.0 / (pos[j] - pos[b])) for b in range(j)]
            fac += [((pos[c] - x - y) / (pos[c] - pos[i] - pos[j]), -1.0 / (pos[c] - pos[i] - pos[j]), -1.0 / (pos[c] - pos[i] - pos[j]))
                    for c in range(i + j + 1, k + 1)]
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
            out.append((gx, gy))
    return out


def gauss_line(order):
    """Gauss-Legendre on [0, 1], order p -> ceil((p + 1) / 2) points."""
    m = (order + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    return [(0.5 * (xi + 1.0), 0.5 * wi) for xi, wi in zip(x, w)]


class Params:
    """Parameters of the coupling operator (test/stokesdarcy2.cc parameter class, ini test/topbottomfiltration.ini:1-20)."""

    def __init__(self, gravity=9.81, alpha=1.0, viscosity=1e-3, porosity=0.4, gamma=1.0, density=1e3, kabs=((1e-7, 0.0), (0.0, 1e-7)), epsilon=1e-2):
        self.gravity, self.alpha, self.viscosity, self.porosity, self.gamma, self.density = gravity, alpha, viscosity, porosity, gamma, density
        self.kabs, self.epsilon = kabs, epsilon


def alpha_coupling(P, stokes_tri, face, darcy_tri, stokes_v, darcy_x, k_v=2, k_d=3, quirk=True):
    """stokes_tri, darcy_tri: three vertex coordinates each, sharing the edge `face` (DUNE numbering) of the Stokes cell;
    stokes_v[d][i]: velocity coefficients (component d, P2 function i); darcy_x[i]: DG-P3 head coefficients.
    Returns (r_stokes_v[d][i], r_darcy[i]).  The pressure block receives nothing (stokesdarcycouplingoperator.hh:219-242)."""
    dim = 2
    vsize, dsize = pk_size(k_v), pk_size(k_d)
    la, lb = FACE_V[face]
    pa, pb = np.asarray(stokes_tri[la], float), np.asarray(stokes_tri[lb], float)
    po = np.asarray(stokes_tri[2 - face], float)
    edge = pb - pa
    ie = math.hypot(edge[0], edge[1])
    n = np.array([edge[1] / ie, -edge[0] / ie])                  # unit outer normal of the inside (Stokes) cell
    if np.dot(n, po - pa) > 0:
        n = -n
    dt = [np.asarray(p, float) for p in darcy_tri]
    ia = [i for i in range(3) if np.allclose(dt[i], pa)][0]
    ib = [i for i in range(3) if np.allclose(dt[i], pb)][0]
    J = np.array([dt[1] - dt[0], dt[2] - dt[0]]).T                # x = p0 + J xi on the Darcy cell
    JinvT = np.linalg.inv(J).T
    qorder = 2 * max(k_v, k_d)                                     # :159-160
    tracePi = sum(P.kabs[i][i] for i in range(dim)) * P.viscosity / P.gravity / P.density      # :172-175
    r_v = [[0.0] * vsize for _ in range(dim)]
    r_d = [0.0] * dsize
    for t, w in gauss_line(qorder):
        pos = pa + t * edge
        sx, sy = REF[la][0] + t * (REF[lb][0] - REF[la][0]), REF[la][1] + t * (REF[lb][1] - REF[la][1])
        dx, dy = REF[ia][0] + t * (REF[ib][0] - REF[ia][0]), REF[ia][1] + t * (REF[ib][1] - REF[ia][1])
        factor = w * ie
        v = pk_basis(k_v, sx, sy)
        psi = pk_basis(k_d, dx, dy)
        js = pk_gradients(k_d, dx, dy)
        gradpsi = [np.zeros(dim) for _ in range(dsize)]            # [UPSTREAM] value-initialised FieldVector entries
        for i in range(vsize if quirk else dsize):                 # :194-195 loops i < vsize
            gradpsi[i] = JinvT @ np.array(js[i])
        phi, gradphi = 0.0, np.zeros(dim)
        for i in range(dsize):
            phi += darcy_x[i] * psi[i]
            gradphi += darcy_x[i] * gradpsi[i]
        u = np.zeros(dim)
        for d in range(dim):
            for i in range(vsize):
                u[d] += stokes_v[d][i] * v[i]
        for i in range(dsize):                                     # :219
            r_d[i] += -P.gamma * P.porosity * float(np.dot(u, n)) * psi[i] * factor
        tang = np.array(P.kabs) @ gradphi
        tang /= P.porosity
        tang += u
        tang -= n * float(np.dot(tang, n))                         # project into the tangential plane
        for d in range(dim):
            for i in range(vsize):
                r_v[d][i] += -P.density * P.gravity * (phi - pos[dim - 1]) * v[i] * n[d] * factor          # :235
            for i in range(vsize):
                r_v[d][i] += P.alpha * math.sqrt(dim) / math.sqrt(tracePi) * tang[d] * v[i] * factor       # :240 Beavers-Joseph
    return r_v, r_d


def pk_interpolate(k, tri, f):
    """Coefficients of the Pk2D Lagrange interpolant of f(x, y) on the triangle `tri`."""
    p = [np.asarray(q, float) for q in tri]
    out = []
    for j in range(k + 1):
        for i in range(k + 1 - j):
            x = p[0] + (p[1] - p[0]) * (i / float(k)) + (p[2] - p[0]) * (j / float(k))
            out.append(f(x[0], x[1]))
    return out


# ---------------------------------------------------------------------------------------------------------------------------------------
# [UPSTREAM] Dune::PDELab::ConvectionDiffusionDG<DarcyParams, DarcyFEM>(SIPG, weightsOn, alpha) (test/stokesdarcy2.cc:733-739) on a
# triangle mesh with DG-Pk spaces (b = 0, c = 0, a piecewise constant tensor A per cell): the published algorithm of dune-pdelab's
# convectiondiffusiondg.hh in the form oracle/oracle.cpp restates for cubes (volume term, interior-penalty skeleton term with harmonic
# weights delta = n.A n and the "Houston" face diameter min(|T_s|, |T_n|) / |F|, weak Dirichlet / Neumann boundary term), now with the
# tensor.  The integrands are polynomials, so any exact rule gives the reference's integrals up to rounding: a conical Gauss product rule
# is used.  Pinned by consistency (a degree-k polynomial solves the scheme exactly), SIPG symmetry and coercivity.
# ---------------------------------------------------------------------------------------------------------------------------------------
def tri_rule_exact(order):
    m = order // 2 + 2
    x, w = np.polynomial.legendre.leggauss(m)
    x, w = 0.5 * (x + 1.0), 0.5 * w
    return [((s, t * (1.0 - s)), ws * wt * (1.0 - s)) for s, ws in zip(x, w) for t, wt in zip(x, w)]


class DGTriangles:
    """residual(x) of ConvectionDiffusionDG on the triangles `conn` of `xy`; DOF = cell * nk + i (cell-blocked, [UPSTREAM] (iii))."""

    def __init__(self, xy, conn, k, A, f, bctype, g, j, alpha=2.0, theta=-1.0, weights=True, intorderadd=0):
        self.xy, self.conn, self.k, self.nk = np.asarray(xy, float), np.asarray(conn), k, pk_size(k)
        self.A = [np.asarray(a, float) for a in A]                  # one 2x2 tensor per cell
        self.f, self.bctype, self.g, self.j = f, bctype, g, j       # f(x,y), bctype(x,y) -> 'D' | 'N' | None at the face centre, g, j
        self.alpha, self.theta, self.weights, self.intorder = alpha, theta, weights, intorderadd + 2 * k
        self.ne = self.conn.shape[0]
        self.ndof = self.ne * self.nk
        edge = {}
        self.nbr = -np.ones((self.ne, 3), dtype=int)
        for e in range(self.ne):
            for fc in range(3):
                a, b = self.conn[e, FACE_V[fc][0]], self.conn[e, FACE_V[fc][1]]
                key = (min(a, b), max(a, b))
                if key in edge:
                    e2, f2 = edge[key]
                    self.nbr[e, fc], self.nbr[e2, f2] = e2, e
                else:
                    edge[key] = (e, fc)

    def _geo(self, e):
        p = self.xy[self.conn[e]]
        J = np.array([p[1] - p[0], p[2] - p[0]]).T
        return p, np.linalg.inv(J).T, abs(np.linalg.det(J))

    def _eval(self, e, lx, ly):
        p, JinvT, ie = self._geo(e)
        phi = np.array(pk_basis(self.k, lx, ly))
        grad = np.array([JinvT @ np.array(gr) for gr in pk_gradients(self.k, lx, ly)])
        return phi, grad, p[0] + (p[1] - p[0]) * lx + (p[2] - p[0]) * ly

    def residual(self, x):
        r = np.zeros(self.ndof)
        nk = self.nk
        for e in range(self.ne):
            xs = x[e * nk:(e + 1) * nk]
            p, JinvT, ie = self._geo(e)
            A_s = self.A[e]
            for (lx, ly), w in tri_rule_exact(self.intorder):                                 # alpha_volume + lambda_volume
                phi, grad, xg = self._eval(e, lx, ly)
                Agradu = A_s @ (xs @ grad)
                factor = w * ie
                r[e * nk:(e + 1) * nk] += (grad @ Agradu - self.f(xg[0], xg[1]) * phi) * factor
            for fc in range(3):
                nb = int(self.nbr[e, fc])
                la, lb = FACE_V[fc]
                pa, pb = p[la], p[lb]
                edge = pb - pa
                fie = math.hypot(edge[0], edge[1])
                n_F = np.array([edge[1], -edge[0]]) / fie
                if np.dot(n_F, p[2 - fc] - pa) > 0:
                    n_F = -n_F
                if nb >= 0 and nb < e:
                    continue                                                                   # interior faces once, from the lower-index cell
                boundary = nb < 0
                An_s = A_s @ n_F
                delta_s = float(An_s @ n_F)
                if boundary:
                    centre = pa + 0.5 * edge
                    bc = self.bctype(centre[0], centre[1])
                    if bc is None:
                        continue
                    omega_s, omega_n, harmonic, vol = 1.0, 0.0, delta_s, 0.5 * ie
                else:
                    pn, JinvTn, ien = self._geo(nb)
                    An_n = self.A[nb] @ n_F
                    delta_n = float(An_n @ n_F)
                    if self.weights:
                        omega_s, omega_n = delta_n / (delta_s + delta_n + 1e-20), delta_s / (delta_s + delta_n + 1e-20)
                        harmonic = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20)
                    else:
                        omega_s = omega_n = 0.5
                        harmonic = 1.0
                    vol = min(0.5 * ie, 0.5 * ien)
                    cn = list(self.conn[nb])
                    ia, ib = cn.index(self.conn[e, la]), cn.index(self.conn[e, lb])
                    xn = x[nb * nk:(nb + 1) * nk]
                if not self.weights and boundary:
                    harmonic = 1.0
                h_F = vol / fie
                penalty = (self.alpha / h_F) * harmonic * self.k * (self.k + 2 - 1)
                for t, w in gauss_line(self.intorder):
                    factor = w * fie
                    lx, ly = REF[la][0] + t * (REF[lb][0] - REF[la][0]), REF[la][1] + t * (REF[lb][1] - REF[la][1])
                    phi_s, grad_s, xg = self._eval(e, lx, ly)
                    u_s, gradu_s = float(xs @ phi_s), xs @ grad_s
                    if boundary and bc == 'N':
                        r[e * nk:(e + 1) * nk] += self.j(xg[0], xg[1]) * phi_s * factor
                        continue
                    if boundary:                                                               # weak Dirichlet (Nitsche)
                        gval = self.g(xg[0], xg[1])
                        rs = -float(An_s @ gradu_s) * factor * phi_s
                        rs += (u_s - gval) * factor * self.theta * (grad_s @ An_s)
                        rs += penalty * (u_s - gval) * factor * phi_s
                        r[e * nk:(e + 1) * nk] += rs
                        continue
                    mx, my = REF[ia][0] + t * (REF[ib][0] - REF[ia][0]), REF[ia][1] + t * (REF[ib][1] - REF[ia][1])
                    phi_n, grad_n, _ = self._eval(nb, mx, my)
                    u_n, gradu_n = float(xn @ phi_n), xn @ grad_n
                    term2 = -(omega_s * float(An_s @ gradu_s) + omega_n * float(An_n @ gradu_n)) * factor
                    term3 = (u_s - u_n) * factor
                    term4 = penalty * (u_s - u_n) * factor
                    r[e * nk:(e + 1) * nk] += term2 * phi_s + term3 * self.theta * omega_s * (grad_s @ An_s) + term4 * phi_s
                    r[nb * nk:(nb + 1) * nk] += -term2 * phi_n + term3 * self.theta * omega_n * (grad_n @ An_n) - term4 * phi_n
        return r

    def jacobian(self):
        """The residual is affine: column j = R(e_j) - R(0) (small meshes only)."""
        r0 = self.residual(np.zeros(self.ndof))
        J = np.zeros((self.ndof, self.ndof))
        for jcol in range(self.ndof):
            ej = np.zeros(self.ndof); ej[jcol] = 1.0
            J[:, jcol] = self.residual(ej) - r0
        return J, r0

    def interpolate(self, u):
        return np.concatenate([pk_interpolate(self.k, self.xy[self.conn[e]], u) for e in range(self.ne)])


# ---------------------------------------------------------------------------------------------------------------------------------------
# [UPSTREAM] Dune::PDELab::TaylorHoodNavierStokes<NavierStokesParams>(params, stokes_q) (test/stokesdarcy2.cc:719-720) on a triangle mesh:
# velocity P2 per component, pressure P1 (test/stokesdarcy2.cc:613-614), composite ordering [v_0 | v_1 | p] (:633-656).  Published
# algorithm of dune-pdelab's taylorhoodnavierstokes.hh (Stokes case, no convection): for every velocity component d and test function i
#     r_v(d, i) += ( mu grad(u_d) . grad(v_i) - p d_d v_i - f_d v_i ) w |det J|        (full-gradient viscous form)
# and for every pressure test function     r_p(i) += -( div u ) q_i w |det J|,
# which makes the operator symmetric.  The SIGN of the continuity rows and the viscous form are assumptions about upstream (flagged);
# the pins are internal: a quadratic divergence-free velocity with a linear pressure solves the scheme exactly, the matrix is symmetric,
# and constants are in the kernel of the pressure gradient block on interior rows.
# Numbering assumption [UPSTREAM (iii)]: per P2 block vertices first (ascending vertex index), then edges; edge indices here are the
# order of first appearance in (cell, local face) order — UG's own edge numbering is an implementation detail of UG.
# ---------------------------------------------------------------------------------------------------------------------------------------
class StokesTaylorHood:
    def __init__(self, xy, conn, mu=1.0, f=lambda x, y: (0.0, 0.0), qorder=4):
        self.xy, self.conn, self.mu, self.f, self.qorder = np.asarray(xy, float), np.asarray(conn), mu, f, qorder
        self.nv, self.ne = self.xy.shape[0], self.conn.shape[0]
        self.edge_index, self.cell_edges = {}, []
        for e in range(self.ne):
            ce = []
            for fc in range(3):
                a, b = self.conn[e, FACE_V[fc][0]], self.conn[e, FACE_V[fc][1]]
                ce.append(self.edge_index.setdefault((min(a, b), max(a, b)), len(self.edge_index)))
            self.cell_edges.append(ce)
        self.n2 = self.nv + len(self.edge_index)                   # size of one P2 block
        self.ndof = 2 * self.n2 + self.nv

    def p2_dofs(self, e):
        """Block-local indices of the six Pk2D (k = 2) functions of cell e: lattice order (0,0),(1/2,0),(1,0),(0,1/2),(1/2,1/2),(0,1):
        vertex 0, edge {0,1} = face 0, vertex 1, edge {0,2} = face 1, edge {1,2} = face 2, vertex 2."""
        c, ed = self.conn[e], self.cell_edges[e]
        return [c[0], self.nv + ed[0], c[1], self.nv + ed[1], self.nv + ed[2], c[2]]

    def dofs(self, e):
        v = self.p2_dofs(e)
        return [v, [self.n2 + i for i in v], [2 * self.n2 + int(i) for i in self.conn[e]]]

    def residual(self, x):
        r = np.zeros(self.ndof)
        for e in range(self.ne):
            p = self.xy[self.conn[e]]
            J = np.array([p[1] - p[0], p[2] - p[0]]).T
            JinvT, ie = np.linalg.inv(J).T, abs(np.linalg.det(J))
            d0, d1, dp = self.dofs(e)
            xv = [x[d0], x[d1]]
            xp = x[dp]
            for (lx, ly), w in tri_rule_exact(self.qorder):
                phi = np.array(pk_basis(2, lx, ly))
                grad = np.array([JinvT @ np.array(g) for g in pk_gradients(2, lx, ly)])
                psi = np.array(pk_basis(1, lx, ly))
                xg = p[0] + (p[1] - p[0]) * lx + (p[2] - p[0]) * ly
                fx = self.f(xg[0], xg[1])
                jacu = [xv[0] @ grad, xv[1] @ grad]                 # grad u_d
                pres = float(xp @ psi)
                divu = jacu[0][0] + jacu[1][1]
                factor = w * ie
                for d, dofs in ((0, d0), (1, d1)):
                    r[dofs] += (self.mu * (grad @ jacu[d]) - pres * grad[:, d] - fx[d] * phi) * factor
                r[dp] += -divu * psi * factor
        return r

    def jacobian(self):
        r0 = self.residual(np.zeros(self.ndof))
        Jm = np.zeros((self.ndof, self.ndof))
        for j in range(self.ndof):
            ej = np.zeros(self.ndof); ej[j] = 1.0
            Jm[:, j] = self.residual(ej) - r0
        return Jm, r0

    def interpolate(self, u, pfun):
        x = np.zeros(self.ndof)
        for e in range(self.ne):
            d0, d1, dp = self.dofs(e)
            tri = self.xy[self.conn[e]]
            x[d0] = pk_interpolate(2, tri, lambda a, b: u(a, b)[0])
            x[d1] = pk_interpolate(2, tri, lambda a, b: u(a, b)[1])
            x[dp] = pk_interpolate(1, tri, pfun)
        return x

    def boundary_velocity_dofs(self):
        """Velocity DOFs on the physical boundary (vertices and edges of faces without a neighbour), both components."""
        count = {}
        for e in range(self.ne):
            for fc in range(3):
                count[self.cell_edges[e][fc]] = count.get(self.cell_edges[e][fc], 0) + 1
        out = set()
        for (a, b), idx in self.edge_index.items():
            if count[idx] == 1:
                out.update([a, b, self.nv + idx])
        return sorted(out) + [self.n2 + i for i in sorted(out)]
