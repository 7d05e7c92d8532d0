"""oracle/um_stokesdarcy.py — TEST INFRASTRUCTURE ONLY.  Groundwork for SURVEY §8 row a12 (no device path yet).

Restates, for one interface edge of a triangle mesh, the reference's Stokes-Darcy coupling operator
StokesDarcyCouplingOperator::alpha_coupling (test/stokesdarcycouplingoperator.hh:104-245) and the local bases it is evaluated
with: Taylor-Hood velocity P2 per component on the Stokes (inside) cell, DG-P3 hydraulic head on the Darcy (outside) cell
(test/stokesdarcy2.cc:613-616, :646-700).  PARITY UNPINNED at the upstream boundary like every oracle in this directory; pinned by
tests/test_oracle_stokesdarcy.py (basis identities, closed-form row sums, hydrostatic equilibrium).

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
