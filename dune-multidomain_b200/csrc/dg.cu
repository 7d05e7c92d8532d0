// dg.cu — QkDG spaces with [UPSTREAM] Dune::PDELab::ConvectionDiffusionDG: BASELINE configs[0] as the reference ships it
// (test/testpoisson-dirichlet-neumann.cc:864-881 dglop(), :883-977 main(): DG-Q2 on both subdomains, SIPG, weightsOn, alpha = 2;
// parameter class ConvectionDiffusionModelProblem :169-285 with A = I, b = 0, c = 0).
//
// What the reference does per traversal (residualassemblerfunctors.hh:16-162, jacobianassemblerfunctors.hh:16-94): alpha_volume /
// lambda_volume per cell; on every interior face whose two cells carry the subproblem alpha_skeleton, once, from the cell with the
// higher index (doSkeletonTwoSided = false, oneSidedDirection = ids > idn, globalassembler.hh:175-183), scattering into BOTH cells'
// rows; on a face whose outside cell does not carry the subproblem — the subdomain interface — alpha_boundary, which returns at
// once because the parameter class answers BCType::None for !is.boundary() (:226-229); on domain boundary faces alpha_boundary
// (weak Dirichlet / Neumann).
//
// Here: owner-computes, no scatter, no atomics.  A QkDG coefficient belongs to one cell, so a matrix row is a list of dense
// (k+1)^d blocks: the cell's own block and one block per face neighbour in the pattern.  On the structured mesh all cells are
// congruent and A is constant, so every block is one of a handful of small matrices that depend only on the face direction:
//   K                        volume (the stiffness matrix: launch_element_matrices, exact.cu)
//   SS_f, SN_f, NS_f, NN_f   jacobian_skeleton of face f of the inside cell
//   BD_f                     jacobian_boundary, Dirichlet branch
// evaluated once per call by k_dg_face_matrices (inside the timed region, like the element matrices).  The Jacobian kernel gives a
// warp to every cell: the rows of a cell are consecutive and share their column set, so the cell's values are one contiguous
// dense image that the lanes write with coalesced stores.  The residual kernel gives a thread to every row and applies it to x
// (the operator is affine, R(x) = J x + b, with b — source, Dirichlet and Neumann data — integrated per row by quadrature).
// Entries agree with the sequential scatter to rounding (1e-12 bar).
#include "exact_common.cuh"

namespace mdb {

#define DG_MAXNL 27
#define DG_MATS 5                                  // SS, SN, NS, NN, BD
#define DG_PHI_OFFSET (6 * DG_MATS * DG_MAXNL * DG_MAXNL)      // phi_i at the volume quadrature points, behind the block matrices
#define DG_SLOT_DOUBLES (DG_PHI_OFFSET + DG_MAXNL * DG_MAXNL)

struct DgSpec {
  int k, nl, st, subdomain;
  int lat_n[3];
  int64_t offset, size;
  const int32_t* lat2dof; const int32_t* dof2lat;
  int cond_kind; uint32_t mask;
  mdb_bctype bc; mdb_function f, g, j;
  double theta, alpha;
  Rule1D rule;                                     // order intorderadd + 2 k, for volume and face terms alike
};

// penalty_factor = (alpha / h_F) * harmonic_average * degree (degree + dim - 1); h_F = min(vol_s, vol_n) / vol_face ("Houston"),
// harmonic_average = 1 for A = I with either weights setting (2 * 1 * 1 / (1 + 1 + 1e-20) rounds to 1; boundary: delta_s = 1)
template <int DIM> __device__ __forceinline__ double dg_penalty(double alpha, int k, const double* ext, int axis)
{
  double vol = ext[0], area = 1.0; bool first = true;
#pragma unroll
  for (int d = 1; d < DIM; ++d) vol *= ext[d];
#pragma unroll
  for (int d = 0; d < DIM; ++d) if (d != axis) { area = first ? ext[d] : area * ext[d]; first = false; }
  const double h_F = vol / area;
  return (alpha / h_F) * 1.0 * k * (k + DIM - 1);
}
template <int DIM> __device__ __forceinline__ double dg_area(const double* ext, int axis)
{
  double area = 1.0; bool first = true;
#pragma unroll
  for (int d = 0; d < DIM; ++d) if (d != axis) { area = first ? ext[d] : area * ext[d]; first = false; }
  return area;
}

// grid: one block per face direction f = 2 axis + side of the inside cell; thread t = i * nl + j owns entry (i, j) of the five
// matrices.  Terms and signs: jacobian_skeleton / jacobian_boundary of ConvectionDiffusionDG with normalflux = 0 (b = 0):
// II diffusion, III consistency (theta = -1 SIPG / +1 NIPG), IV interior penalty; omega_s = omega_n = 1/2 (A = I).
template <int DIM> __global__ void __launch_bounds__(768) k_dg_face_matrices(int k, double theta, double alpha, Rule1D R, double ext0, double ext1, double ext2,
                                                                             double weight, double* T)
{
  const int f = blockIdx.x, axis = f / 2, side = f % 2;
  const int nl = ipow(k + 1, DIM);
  const int t = threadIdx.x;
  const double ext[3] = {ext0, ext1, ext2};
  const double jit = 1.0 / ext[axis], nrm = side ? 1.0 : -1.0;
  const double area = dg_area<DIM>(ext, axis), penalty = dg_penalty<DIM>(alpha, k, ext, axis);
  const double omega = 0.5;
  const int nq = ipow(R.m, DIM - 1);
  // shape functions of both cells at the face points, evaluated once per (point, function) by the first nq * nl threads
  __shared__ double s_ps[9][DG_MAXNL], s_pn[9][DG_MAXNL], s_as[9][DG_MAXNL], s_an[9][DG_MAXNL], s_w[9];
  for (int u = t; u < nq * nl; u += blockDim.x) {
    const int q = u / nl, ii = u % nl;
    double pos[3], w; quad_point<DIM - 1>(R, q, pos, w);
    double ls[3], ln[3]; int tt = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { if (d == axis) { ls[d] = (double)side; ln[d] = (double)(1 - side); } else { ls[d] = pos[tt]; ln[d] = pos[tt]; ++tt; } }
    double gr[3];
    s_ps[q][ii] = qk_phi<DIM>(k, ii, ls); s_pn[q][ii] = qk_phi<DIM>(k, ii, ln);
    qk_grad<DIM>(k, ii, ls, gr); s_as[q][ii] = nrm * (jit * gr[axis]);      // An_F_s * tgradphi_s[i] = n . grad phi_i (inside cell)
    qk_grad<DIM>(k, ii, ln, gr); s_an[q][ii] = nrm * (jit * gr[axis]);      // ... of the outside cell, with the INSIDE normal
    if (ii == 0) s_w[q] = w;
  }
  if (f == 0 && ipow(R.m, DIM) <= DG_MAXNL) {                 // PHI[q * nl + i] for the residual's source term (rule of order 2k: nl points)
    const int nqv = ipow(R.m, DIM);
    for (int u = t; u < nqv * nl; u += blockDim.x) {
      double xl[3], w; quad_point<DIM>(R, u / nl, xl, w);
      T[DG_PHI_OFFSET + u] = qk_phi<DIM>(k, u % nl, xl);
    }
  }
  __syncthreads();
  if (t >= nl * nl) return;
  const int i = t / nl, j = t % nl;
  double ss = 0.0, sn = 0.0, ns = 0.0, nn = 0.0, bd = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double ps_i = s_ps[q][i], ps_j = s_ps[q][j], pn_i = s_pn[q][i], pn_j = s_pn[q][j];
    const double as_i = s_as[q][i], as_j = s_as[q][j], an_i = s_an[q][i], an_j = s_an[q][j];
    const double factor = s_w[q] * area, ipf = penalty * factor;
    ss += -as_j * omega * factor * ps_i + ps_j * factor * theta * omega * as_i + ps_j * ipf * ps_i;
    sn += -an_j * omega * factor * ps_i - pn_j * factor * theta * omega * as_i - pn_j * ipf * ps_i;
    ns += as_j * omega * factor * pn_i + ps_j * factor * theta * omega * an_i - ps_j * ipf * pn_i;
    nn += an_j * omega * factor * pn_i - pn_j * factor * theta * omega * an_i + pn_j * ipf * pn_i;
    bd += -as_j * factor * ps_i + ps_j * factor * theta * as_i + ps_j * ipf * ps_i;
  }
  double* o = T + (size_t)f * DG_MATS * DG_MAXNL * DG_MAXNL + t;
  o[0 * DG_MAXNL * DG_MAXNL] = weight * ss; o[1 * DG_MAXNL * DG_MAXNL] = weight * sn; o[2 * DG_MAXNL * DG_MAXNL] = weight * ns;
  o[3 * DG_MAXNL * DG_MAXNL] = weight * nn; o[4 * DG_MAXNL * DG_MAXNL] = weight * bd;
}

// what a row needs to know about its cell: neighbours, their role, and where their blocks start
template <int DIM> struct DgRow {
  int ijk[3], li; int64_t cell; bool active;
  int32_t first_own;
  // per face: 0 nothing, 1 skeleton seen from this cell as inside, 2 skeleton seen as outside, 3 Dirichlet boundary, 4 Neumann boundary
  int kind[2 * DIM]; int32_t first_nb[2 * DIM];
  __device__ void init(const MeshDesc& m, const uint8_t* __restrict__ sd, const DgSpec& S, int64_t gi)
  {
    // 32-bit arithmetic throughout (a child has < 2^31 lattice points): 64-bit divisions cost ~100 instructions each and every
    // lane of a warp-per-cell kernel runs this
    const unsigned r = (unsigned)S.dof2lat[gi], n0 = (unsigned)S.lat_n[0], n1 = (unsigned)S.lat_n[1], st = (unsigned)S.st;
    const unsigned r1 = r / n0;
    unsigned p[3]; p[0] = r - r1 * n0; p[2] = (DIM > 2) ? r1 / n1 : 0u; p[1] = r1 - p[2] * n1;
    int a[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (d < DIM) { const unsigned c = p[d] / st; ijk[d] = (int)c; a[d] = (int)(p[d] - c * st); } else { ijk[d] = 0; a[d] = 0; }
    }
    li = a[0] + (S.k + 1) * (a[1] + (S.k + 1) * a[2]);
    cell = ijk[0] + (int64_t)m.n[0] * (ijk[1] + (int64_t)m.n[1] * ijk[2]);
    active = cond_applies(S.cond_kind, S.mask, sd[cell]);
    first_own = (int32_t)(S.offset + gi - li);
    for (int f = 0; f < 2 * DIM; ++f) {
      kind[f] = 0; first_nb[f] = -1;
      if (!active) continue;
      const int axis = f / 2, side = f % 2;
      int n3[3] = {ijk[0], ijk[1], ijk[2]};
      n3[axis] += side ? 1 : -1;
      if (n3[axis] < 0 || n3[axis] >= m.n[axis]) {
        // domain boundary: boundary type at the face centre
        double xg[3];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          const double lo = m.coord(d, ijk[d]), up = m.coord(d, ijk[d] + 1);
          xg[d] = (d == axis) ? (side ? up : lo) : lo + 0.5 * (up - lo);
        }
        const int bc = eval_bctype(S.bc, DIM, xg, true);
        kind[f] = bc == 0 ? 3 : (bc == 1 ? 4 : 0);
        continue;
      }
      const int64_t nb = n3[0] + (int64_t)m.n[0] * (n3[1] + (int64_t)m.n[1] * n3[2]);
      const uint8_t sn = sd[nb];
      if (!(cond_applies(S.cond_kind, S.mask, sn) && (S.subdomain < 0 || ((sn >> S.subdomain) & 1)))) continue;   // interface: alpha_boundary, BCType::None
      kind[f] = cell > nb ? 1 : 2;
      const int64_t q = (int64_t)S.st * n3[0] + (int64_t)S.lat_n[0] * ((int64_t)S.st * n3[1] + (int64_t)S.lat_n[1] * ((int64_t)S.st * n3[2]));
      first_nb[f] = (int32_t)(S.offset + S.lat2dof[q]);
    }
  }
  // own-block row and, per face, the neighbour-block row (pointers into the tables; nullptr = no block)
  __device__ void own_row(const DgSpec& S, const double* __restrict__ K, const double* __restrict__ T, double* acc) const
  {
    for (int jj = 0; jj < S.nl; ++jj) acc[jj] = K[li * S.nl + jj];
    for (int f = 0; f < 2 * DIM; ++f) {
      const double* M = nullptr;
      if (kind[f] == 1) M = T + ((size_t)f * DG_MATS + 0) * DG_MAXNL * DG_MAXNL;              // SS_f
      else if (kind[f] == 2) M = T + ((size_t)(f ^ 1) * DG_MATS + 3) * DG_MAXNL * DG_MAXNL;   // NN of the neighbour's face f^1
      else if (kind[f] == 3) M = T + ((size_t)f * DG_MATS + 4) * DG_MAXNL * DG_MAXNL;         // BD_f
      if (M) for (int jj = 0; jj < S.nl; ++jj) acc[jj] += M[li * S.nl + jj];
    }
  }
  __device__ const double* nb_row(const DgSpec& S, const double* __restrict__ T, int f) const
  {
    if (kind[f] == 1) return T + ((size_t)f * DG_MATS + 1) * DG_MAXNL * DG_MAXNL + li * S.nl;        // SN_f
    if (kind[f] == 2) return T + ((size_t)(f ^ 1) * DG_MATS + 2) * DG_MAXNL * DG_MAXNL + li * S.nl;  // NS of the neighbour's face
    return nullptr;
  }
};

// The (k+1)^d rows of a cell are consecutive DOFs with the same column set, so a cell's values are one contiguous
// nl x (nblk nl) row-major image in the value array.  A block of the image is the cell's own one (K + the face terms seen from this
// cell), a skeleton neighbour's (SN_f / NS_f^1), or a block that only a coupling operator fills (zero here; the coupling kernels add
// to it afterwards on the same stream).  Which is which is a 42-bit SIGNATURE of the cell (face kinds, block roles, block count), and
// on a structured mesh long runs of consecutive cells share it.  A warp takes 32 consecutive cells: each LANE works out the
// signature of one cell (the index arithmetic and the dependent loads are spent once per cell, not once per lane and cell), then
// the warp walks through the 32 cells, rebuilds its image in shared memory only when the signature changes, and copies it out
// with coalesced stores — a few instructions per 256 bytes written.
template <int DIM, int KDEG>
__global__ void __launch_bounds__(256) k_dg_jacobian(MeshDesc m, const uint8_t* __restrict__ sd, DgSpec S, const double* __restrict__ K,
                                                     const double* __restrict__ T, const int64_t* __restrict__ rowptr,
                                                     const int32_t* __restrict__ colidx, double* __restrict__ val, int overwrite, int cpw)
{
  constexpr int NL = cpow(KDEG + 1, DIM), NF = 2 * DIM, MM = DG_MAXNL * DG_MAXNL, IMG = NL * (NF + 1) * NL;
  extern __shared__ double s_img[];
  double* img = s_img + (size_t)(threadIdx.x >> 5) * IMG;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ncell = S.size / NL;
  unsigned long long cur = ~0ull;                             // signature of the image in shared memory
  for (int64_t b0 = warp * cpw; b0 < ncell; b0 += nwarps * cpw) {       // cpw <= 32 cells per warp and pass (fewer on small meshes)
    // ---- one cell per lane: signature and position
    const int64_t b = b0 + lane;
    unsigned long long sig = ~0ull; long long rs = 0; int len = 0, act = 0;
    if (lane < cpw && b < ncell) {
      const int64_t gi = b * NL, g = S.offset + gi;
      rs = rowptr[g]; len = (int)(rowptr[g + 1] - rs);
      DgRow<DIM> R; R.init(m, sd, S, gi);                     // li = 0: the cell's first row
      act = R.active ? 1 : 0;
      if (act) {
        const int nblk = len / NL;
        sig = (unsigned long long)nblk << 39;
#pragma unroll
        for (int f = 0; f < NF; ++f) sig |= (unsigned long long)R.kind[f] << (3 * f);
        for (int bb = 0; bb < nblk; ++bb) {                   // role of block bb: face f (0..5), 6 = own, 7 = filled by a coupling only
          const int32_t first = colidx[rs + bb * NL];
          int role = 7;
          if (first == R.first_own) role = 6;
          else {
#pragma unroll
            for (int f = 0; f < NF; ++f) if (first == R.first_nb[f] && (R.kind[f] == 1 || R.kind[f] == 2)) role = f;
          }
          sig |= (unsigned long long)role << (18 + 3 * bb);
        }
      }
    }
    // ---- the warp walks through its cells
    const int ncl = (ncell - b0 < cpw) ? (int)(ncell - b0) : cpw;
    for (int l = 0; l < ncl; ++l) {
      const unsigned long long sg = __shfl_sync(0xffffffffu, sig, l);
      const long long rsl = __shfl_sync(0xffffffffu, rs, l);
      const int lenl = __shfl_sync(0xffffffffu, len, l), actl = __shfl_sync(0xffffffffu, act, l);
      const int total = lenl * NL;
      double* dst = val + rsl;
      if (!actl) { if (overwrite) for (int e = lane; e < total; e += 32) dst[e] = 0.0; continue; }
      if (sg != cur) {                                        // (re)build the image: rare
        __syncwarp();
        for (int e = lane; e < total; e += 32) {
          const int i = e / lenl, rem = e - i * lenl, bb = rem / NL, j = rem - bb * NL, ij = i * NL + j;
          const int role = (int)((sg >> (18 + 3 * bb)) & 7);
          double v = 0.0;
          if (role == 6) {
            v = K[ij];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
              const int kd = (int)((sg >> (3 * f)) & 7);
              if (kd == 1) v += T[(f * DG_MATS + 0) * MM + ij];               // SS_f
              else if (kd == 2) v += T[((f ^ 1) * DG_MATS + 3) * MM + ij];    // NN of the neighbour's face f^1
              else if (kd == 3) v += T[(f * DG_MATS + 4) * MM + ij];          // BD_f
            }
          } else if (role < 6) {
            const int kd = (int)((sg >> (3 * role)) & 7);
            v = kd == 1 ? T[(role * DG_MATS + 1) * MM + ij] : T[((role ^ 1) * DG_MATS + 2) * MM + ij];     // SN_f / NS of the neighbour's face
          }
          img[e] = v;
        }
        cur = sg;
        __syncwarp();
      }
      if (overwrite) { for (int e = lane; e < total; e += 32) dst[e] = img[e]; }
      else { for (int e = lane; e < total; e += 32) { const double v = img[e]; if (v != 0.0) dst[e] += v; } }
    }
  }
}

// R(x)_row = (J x)_row + b_row:  b = lambda_volume (- int f phi_i) + on Dirichlet faces the g terms of alpha_boundary
// (- theta g (n . grad phi_i) - penalty g phi_i) + on Neumann faces int j phi_i.
// One thread per row; the block size is a multiple of the rows per cell, so the threads of a cell sit in one block: the volume rule
// of order 2k has as many points as the cell has rows, thread li evaluates the source at point li (once per cell, not once per row)
// and the cell's threads read each other's values from shared memory; phi_i at the volume points comes from a table (PHI, written
// by k_dg_face_matrices) instead of nine Lagrange evaluations with their divisions per row.
template <int DIM>
__global__ void __launch_bounds__(128) k_dg_residual(MeshDesc m, const uint8_t* __restrict__ sd, DgSpec S, const double* __restrict__ K,
                                                     const double* __restrict__ T, double weight, double time, const double* __restrict__ x,
                                                     double* __restrict__ res)
{
  __shared__ double s_f[128];
  const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool in = gi < S.size;
  DgRow<DIM> R; R.init(m, sd, S, in ? gi : 0);
  const bool act = in && R.active;
  const int nl = S.nl, li = R.li;
  double lo[3], ext[3];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { lo[d] = m.coord(d, R.ijk[d]); ext[d] = m.coord(d, R.ijk[d] + 1) - lo[d]; }
  const int nqv = ipow(S.rule.m, DIM);
  const bool shared_f = S.f.kind != MDB_FN_ZERO && nqv == nl;      // uniform over the grid
  if (shared_f) {
    double fq = 0.0;
    if (act) {
      double detj = ext[0];
#pragma unroll
      for (int d = 1; d < DIM; ++d) detj *= ext[d];
      double xl[3], w, xg[3]; quad_point<DIM>(S.rule, li, xl, w);
#pragma unroll
      for (int d = 0; d < DIM; ++d) xg[d] = lo[d] + xl[d] * ext[d];
      fq = eval_function(S.f, DIM, xg, time) * (w * detj);
    }
    s_f[threadIdx.x] = fq;
    __syncthreads();
  }
  if (!act) return;
  double sum = 0.0;
  {
    const double* Kr = K + li * nl;
    const double* Mr[2 * DIM];
#pragma unroll
    for (int f = 0; f < 2 * DIM; ++f) {
      Mr[f] = nullptr;
      if (R.kind[f] == 1) Mr[f] = T + ((size_t)f * DG_MATS + 0) * DG_MAXNL * DG_MAXNL + li * nl;              // SS_f
      else if (R.kind[f] == 2) Mr[f] = T + ((size_t)(f ^ 1) * DG_MATS + 3) * DG_MAXNL * DG_MAXNL + li * nl;   // NN of the neighbour's face
      else if (R.kind[f] == 3) Mr[f] = T + ((size_t)f * DG_MATS + 4) * DG_MAXNL * DG_MAXNL + li * nl;         // BD_f
    }
    const double* xo = x + R.first_own;
    for (int jj = 0; jj < nl; ++jj) {
      double v = Kr[jj];
#pragma unroll
      for (int f = 0; f < 2 * DIM; ++f) if (Mr[f]) v += Mr[f][jj];
      sum += v * xo[jj];
    }
  }
  for (int f = 0; f < 2 * DIM; ++f) {
    const double* M = R.nb_row(S, T, f);
    if (!M) continue;
    const double* xn = x + R.first_nb[f];
    double s2 = 0.0;
    for (int jj = 0; jj < nl; ++jj) s2 += M[jj] * xn[jj];
    sum += s2;
  }
  double b = 0.0;
  if (shared_f) {
    const double* PHI = T + DG_PHI_OFFSET;                          // PHI[q * nl + i] = phi_i at volume point q
    const double* fc = s_f + (threadIdx.x - li);
    for (int q = 0; q < nl; ++q) b -= fc[q] * PHI[q * nl + li];
  } else if (S.f.kind != MDB_FN_ZERO) {
    double detj = ext[0];
#pragma unroll
    for (int d = 1; d < DIM; ++d) detj *= ext[d];
    for (int q = 0; q < nqv; ++q) {
      double xl[3], w, xg[3]; quad_point<DIM>(S.rule, q, xl, w);
#pragma unroll
      for (int d = 0; d < DIM; ++d) xg[d] = lo[d] + xl[d] * ext[d];
      b += -eval_function(S.f, DIM, xg, time) * qk_phi<DIM>(S.k, li, xl) * (w * detj);
    }
  }
  for (int f = 0; f < 2 * DIM; ++f) {
    if (R.kind[f] < 3) continue;
    const int axis = f / 2, side = f % 2;
    const double area = dg_area<DIM>(ext, axis), penalty = dg_penalty<DIM>(S.alpha, S.k, ext, axis);
    const double jit = 1.0 / ext[axis], nrm = side ? 1.0 : -1.0;
    const int nq = ipow(S.rule.m, DIM - 1);
    for (int q = 0; q < nq; ++q) {
      double pos[3], w; quad_point<DIM - 1>(S.rule, q, pos, w);
      double ls[3], xg[3]; int tt = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) { ls[d] = (d == axis) ? (double)side : pos[tt++]; xg[d] = lo[d] + ls[d] * ext[d]; }
      const double phi = qk_phi<DIM>(S.k, li, ls), factor = w * area;
      if (R.kind[f] == 4) { b += eval_function(S.j, DIM, xg, time) * phi * factor; continue; }
      double gr[3]; qk_grad<DIM>(S.k, li, ls, gr);
      const double gv = eval_function(S.g, DIM, xg, time);
      b += -gv * factor * S.theta * (nrm * (jit * gr[axis])) - penalty * gv * factor * phi;
    }
  }
  res[S.offset + gi] += sum + weight * b;
}

static DgSpec dg_spec(Ctx* c, const GridOp& go, int child, int* slot)
{
  const Child& ch = c->children[child];
  DgSpec S; std::memset(&S, 0, sizeof(S));
  S.k = ch.k; S.nl = ch.nloc(c->mesh.dim); S.st = ch.stride(); S.subdomain = ch.subdomain;
  for (int d = 0; d < 3; ++d) S.lat_n[d] = ch.lat_n[d];
  S.offset = ch.offset; S.size = ch.size; S.lat2dof = ch.d_lat2dof; S.dof2lat = ch.d_dof2lat;
  *slot = -1;
  for (size_t i = 0; i < go.sps.size(); ++i) {
    const SubProblemDesc& sp = c->sps[go.sps[i]];
    if (sp.children[0] != child) continue;
    MDB_REQUIRE(*slot < 0, MDB_EINVAL, "QkDG spaces: one subproblem per child space and grid operator");
    MDB_REQUIRE(sp.op_id == MDB_OP_CONVDIFF_DG, MDB_EINVAL, "QkDG spaces carry ConvectionDiffusionDG");
    *slot = (int)i;
    S.cond_kind = sp.cond_kind; S.mask = sp.mask; S.bc = sp.cddg.bc; S.f = sp.cddg.f; S.g = sp.cddg.g; S.j = sp.cddg.j;
    S.theta = sp.cddg.method == MDB_DG_METHOD_SIPG ? -1.0 : 1.0; S.alpha = sp.cddg.alpha;
    S.rule = gauss1d(sp.quad_order());
  }
  return S;
}

static double* dg_tables(Ctx* c, const DgSpec& S, int slot, double weight)
{
  MDB_REQUIRE(S.rule.m <= 3, MDB_EINVAL, "ConvectionDiffusionDG: face rule larger than the block-table kernel's staging (intorderadd too large)");
  if (!c->d_dgtab) c->d_dgtab = dalloc<double>((size_t)MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN * DG_SLOT_DOUBLES);
  double* T = c->d_dgtab + (size_t)slot * DG_SLOT_DOUBLES;
  const MeshDesc& m = c->mesh;
  double ext[3] = {1, 1, 1};
  for (int d = 0; d < m.dim; ++d) ext[d] = m.coord(d, 1) - m.coord(d, 0);
  const int threads = ((S.nl * S.nl + 31) / 32) * 32;
  if (m.dim == 2) MDB_LAUNCH(c, k_dg_face_matrices<2>, 4, threads, 0, S.k, S.theta, S.alpha, S.rule, ext[0], ext[1], ext[2], weight, T);
  else MDB_LAUNCH(c, k_dg_face_matrices<3>, 6, threads, 0, S.k, S.theta, S.alpha, S.rule, ext[0], ext[1], ext[2], weight, T);
  return T;
}

// launch_element_matrices must have run (K of slot `slot` = the stiffness matrix times weight)
void jacobian_dg(Ctx* c, const GridOp& go, Matrix& M, int child, double weight, bool overwrite)
{
  const MeshDesc& m = c->mesh;
  int slot; DgSpec S = dg_spec(c, go, child, &slot);
  if (S.size == 0) return;
  if (slot < 0) {
    if (overwrite) MDB_CUDA(cudaMemsetAsync(M.d_val + M.child_begin[child], 0, (M.child_end[child] - M.child_begin[child]) * sizeof(double), c->stream));
    return;
  }
  const double* T = dg_tables(c, S, slot, weight);
  const double* K = c->d_Ae + (size_t)slot * MDB_MAXLOC * MDB_MAXLOC;
  const int64_t ncell = S.size / S.nl;
  const int nf = 2 * m.dim;
  const size_t img_bytes = (size_t)S.nl * (nf + 1) * S.nl * sizeof(double);     // one image per warp
  int wpb = (int)(48 * 1024 / img_bytes);                                       // warps per block
  if (wpb > 8) wpb = 8;
  MDB_REQUIRE(wpb >= 1, MDB_EINVAL, "internal: DG cell image larger than shared memory");
  // cells per warp and pass.  Measured at 131 072 cells per child (512^2 DG-Q2): 32 cells (4096 warps) 108 us, 16 cells (8192 warps)
  // 136-148 us — fewer image rebuilds beat more warps; smaller batches only when the mesh cannot give every SM a few warps
  int cpw = 32;
  while (cpw > 4 && ncell / cpw < 148 * 16) cpw /= 2;
  static const int wpb_env = getenv("MDB_DG_WPB") ? atoi(getenv("MDB_DG_WPB")) : 0;   // A/B switch for profiling only
  if (wpb_env > 0 && wpb_env < wpb) wpb = wpb_env;
  int grid = cdiv(ncell, cpw * wpb);
  const int cap = 148 * (64 / wpb);
  if (grid > cap) grid = cap;
  const size_t smem = img_bytes * wpb;
  if (m.dim == 2 && S.k == 1) MDB_LAUNCH(c, (k_dg_jacobian<2, 1>), grid, 32 * wpb, smem, m, c->d_cell_sd, S, K, T, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, cpw);
  else if (m.dim == 2) MDB_LAUNCH(c, (k_dg_jacobian<2, 2>), grid, 32 * wpb, smem, m, c->d_cell_sd, S, K, T, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, cpw);
  else if (S.k == 1) MDB_LAUNCH(c, (k_dg_jacobian<3, 1>), grid, 32 * wpb, smem, m, c->d_cell_sd, S, K, T, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, cpw);
  else MDB_LAUNCH(c, (k_dg_jacobian<3, 2>), grid, 32 * wpb, smem, m, c->d_cell_sd, S, K, T, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, cpw);
}

void residual_dg(Ctx* c, const GridOp& go, int child, double weight, const double* d_x, double* d_r)
{
  const MeshDesc& m = c->mesh;
  int slot; DgSpec S = dg_spec(c, go, child, &slot);
  if (S.size == 0 || slot < 0) return;
  const double* T = dg_tables(c, S, slot, weight);
  const double* K = c->d_Ae + (size_t)slot * MDB_MAXLOC * MDB_MAXLOC;
  const int bs = S.nl * (128 / S.nl);                          // a multiple of the rows per cell: the threads of a cell share a block
  if (m.dim == 2) MDB_LAUNCH(c, k_dg_residual<2>, cdiv(S.size, bs), bs, 0, m, c->d_cell_sd, S, K, T, weight, c->time, d_x, d_r);
  else MDB_LAUNCH(c, k_dg_residual<3>, cdiv(S.size, bs), bs, 0, m, c->d_cell_sd, S, K, T, weight, c->time, d_x, d_r);
}

} // namespace mdb
