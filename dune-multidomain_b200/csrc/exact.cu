// exact.cu — kernels whose arithmetic follows the reference's operation order term by term.
// COMPILED WITH --fmad=false (see Makefile): no FMA contraction, so that the finite-difference coupling Jacobians
// (NumericalJacobianCoupling, couplingutilities.hh:59-141, eps = 1e-8) reproduce the CPU evaluation; a one-ulp change in a
// residual evaluation is amplified to ~1e-8 in (up-down)/delta.
//
//   k_element_matrices  local operator tables: jacobian_volume of Poisson ([UPSTREAM] Laplace::jacobian_volume) and of
//                       L2 / test/simpletimeoperator.hh:68-101 on the (congruent) cells of a structured mesh
//   k_coupling_residual alpha_coupling of ContinuousValueContinuousFlowCoupling (test/proportionalflowcoupling.hh:113-233)
//                       and ProportionalFlowCoupling (:27-86), one thread per interface face
//   k_coupling_jacobian_staged  jacobian_coupling: (face, evaluation) threads; evaluation 0 = base line, 1+j = perturbation j
//   k_lambda_volume / k_lambda_boundary   source and Neumann terms of Poisson (test/instationarypoissonoperator.hh:103-167)
#include "exact_common.cuh"

namespace mdb {

// ---- element matrices --------------------------------------------------------------------------------
struct AeSpec { int n; int op[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN]; int k[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN];
                Rule1D rule[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN]; };

// grid: one block per slot; thread t = i*nloc + j computes entry (i,j) with the reference's summation order over the
// quadrature points.  ext[d] = up-lo of the (congruent) cells.  The shape functions are evaluated once per (quadrature
// point, i) — in parallel over the block, chunk by chunk — and parked in shared memory; the per-entry sums then run over
// the points in rule order, so the result has the same bits as evaluating everything per entry.
#define AE_QCHUNK 32
template <int DIM> __global__ void __launch_bounds__(768) k_element_matrices(AeSpec S, double ext0, double ext1, double ext2, double weight, double* Ae)
{
  __shared__ double sb[AE_QCHUNK][MDB_MAXLOC][4];     // {phi, dphi/dx0, dphi/dx1, dphi/dx2} in reference coordinates
  __shared__ double sw[AE_QCHUNK];
  const int slot = blockIdx.x;
  const int k = S.k[slot];
  const int nloc = ipow(k + 1, DIM);
  const int t = threadIdx.x;
  const bool active = t < nloc * nloc;
  const int i = active ? t / nloc : 0, j = active ? t % nloc : 0;
  const double ext[3] = {ext0, ext1, ext2};
  double jit[3];
#pragma unroll
  for (int d = 0; d < DIM; ++d) jit[d] = 1.0 / ext[d];
  double detj = ext[0];
#pragma unroll
  for (int d = 1; d < DIM; ++d) detj *= ext[d];
  const Rule1D& R = S.rule[slot];
  const int nq = ipow(R.m, DIM);
  double a = 0.0;
  for (int q0 = 0; q0 < nq; q0 += AE_QCHUNK) {
    const int nc = (nq - q0 < AE_QCHUNK) ? nq - q0 : AE_QCHUNK;
    __syncthreads();
    for (int e = t; e < nc * nloc; e += blockDim.x) {
      const int ql = e / nloc, ii = e % nloc;
      double x[3], w; quad_point<DIM>(R, q0 + ql, x, w);
      double g[3] = {0.0, 0.0, 0.0};
      qk_grad<DIM>(k, ii, x, g);
      sb[ql][ii][0] = qk_phi<DIM>(k, ii, x); sb[ql][ii][1] = g[0]; sb[ql][ii][2] = g[1]; sb[ql][ii][3] = g[2];
      if (ii == 0) sw[ql] = w;
    }
    __syncthreads();
    if (active) {
      for (int ql = 0; ql < nc; ++ql) {
        const double factor = sw[ql] * detj;
        if (S.op[slot] == MDB_OP_POISSON) {
          double dot = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) { double a_i = 0.0 + jit[d] * sb[ql][i][1 + d]; double a_j = 0.0 + jit[d] * sb[ql][j][1 + d]; dot += a_j * a_i; }
          a += weight * (dot * factor);
        } else {   // MDB_OP_L2
          a += weight * (sb[ql][j][0] * sb[ql][i][0] * factor);
        }
      }
    }
  }
  double* A = Ae + slot * MDB_MAXLOC * MDB_MAXLOC;
  if (active) A[i * nloc + j] = a;
  __syncthreads();
  // Q1 only: the row stencil of a vertex all of whose 2^DIM incident cells carry this operator, summed over the cells in
  // ascending cell index exactly like the per-row gather of assemble.cu (stencil_row); stored behind the element matrix.
  constexpr int NS = (DIM == 3) ? 27 : 9;
  if (k == 1 && t < NS) {
    double s = 0.0;
    for (int cz = (DIM > 2 ? -1 : 0); cz <= 0; ++cz)
      for (int cy = -1; cy <= 0; ++cy)
        for (int cx = -1; cx <= 0; ++cx) {
          const int li = (-cx) | ((-cy) << 1) | ((DIM > 2 ? -cz : 0) << 2);
          for (int lj = 0; lj < nloc; ++lj) {
            const int bx = lj & 1, by = (lj >> 1) & 1, bz = (DIM > 2) ? (lj >> 2) & 1 : 0;
            const int o = (DIM > 2 ? (cz + bz + 1) * 9 : 0) + (cy + by + 1) * 3 + (cx + bx + 1);
            if (o == t) s += A[li * nloc + lj];
          }
        }
    A[nloc * nloc + t] = s;
  }
}

void launch_element_matrices(Ctx* c, const std::vector<int>& sps, double weight)
{
  const MeshDesc& m = c->mesh;
  AeSpec S; S.n = (int)sps.size();
  MDB_REQUIRE(S.n <= MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN, MDB_EINVAL, "too many subproblems in one grid operator");
  if (S.n == 0) return;
  for (int i = 0; i < S.n; ++i) {
    const SubProblemDesc& sp = c->sps[sps[i]];
    // ConvectionDiffusionDG with A = I, b = 0, c = 0: its volume matrix is the stiffness matrix under the rule of order intorderadd + 2 k
    S.op[i] = sp.op_id == MDB_OP_CONVDIFF_DG ? MDB_OP_POISSON : sp.op_id; S.k[i] = c->children[sp.children[0]].k; S.rule[i] = gauss1d(sp.quad_order());
  }
  double ext[3] = {1, 1, 1};
  // the extent of the (congruent) cells: coord(1) - coord(0) evaluated at global offset 0, so that every rank of a partition gets
  // the bits of the single-rank matrix (with off > 0 the difference of two large products differs in the last place)
  for (int d = 0; d < m.dim; ++d) ext[d] = (m.glower[d] + 1.0 * m.h[d]) - (m.glower[d] + 0.0 * m.h[d]);
  int nthreads = 256;      // at least 256: the shape-function tables of a quadrature chunk are evaluated by the whole block
  for (int i = 0; i < S.n; ++i) { int nl = 1; for (int d = 0; d < m.dim; ++d) nl *= (S.k[i] + 1); int t = ((nl * nl + 31) / 32) * 32; if (t > nthreads) nthreads = t; }
  if (m.dim == 2) MDB_LAUNCH(c, k_element_matrices<2>, S.n, nthreads, 0, S, ext[0], ext[1], ext[2], weight, c->d_Ae);
  else MDB_LAUNCH(c, k_element_matrices<3>, S.n, nthreads, 0, S, ext[0], ext[1], ext[2], weight, c->d_Ae);
}

// Shape-function tables on the faces of the reference cell.  The values phi_i(local) and d phi_i/d local_d at the face
// quadrature points depend only on (face axis, which side of the cell, quadrature point), not on the cell, so they are
// evaluated once — with the very same qk_phi / qk_grad arithmetic, hence bit-identical to evaluating them per face as
// the reference does — and looked up by every face.  Layout: tab[(((axis*2+sideval)*nq + q)*NL + i)*4 + {phi, d0, d1, d2}].
#define MDB_FTAB_DOUBLES (6 * 25 * 8 * 4)
template <int DIM> __global__ void k_face_tables(Rule1D R, double* tab)
{
  constexpr int NL = 1 << DIM;
  const int nq = ipow(R.m, DIM - 1);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * DIM * nq * NL) return;
  const int i = t % NL, q = (t / NL) % nq, c = t / (NL * nq);
  const int axis = c / 2, sideval = c % 2;
  double pos[3], w; quad_point<DIM - 1>(R, q, pos, w);
  double local[3]; int tt = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) local[d] = (d == axis) ? (double)sideval : pos[tt++];
  double js[3] = {0.0, 0.0, 0.0};
  qk_grad<DIM>(1, i, local, js);
  double* o = tab + (size_t)t * 4;
  o[0] = qk_phi<DIM>(1, i, local); o[1] = js[0]; o[2] = js[1]; o[3] = js[2];
}

// alpha_coupling restated for Q1 on both sides (NL = 2^DIM shape functions per side); r_s, r_n accumulate weight*value.
template <int DIM>
__device__ void alpha_coupling(const CouplingParams& P, const FaceGeo<DIM>& G, const double* __restrict__ tab, const double* x1,
                               const double* x2, double* r1, double* r2, double weight)
{
  constexpr int NL = 1 << DIM;
  const int nq = ipow(P.rule.m, DIM - 1);
  const double h_F = G.cornerDistance01();
  const double* tab1 = tab + (size_t)((G.axis * 2 + G.side) * nq) * NL * 4;          // geometryInInside: local[axis] = side
  const double* tab2 = tab + (size_t)((G.axis * 2 + (1 - G.side)) * nq) * NL * 4;    // geometryInOutside: 1 - side
  double jit1[DIM], jit2[DIM];     // jacobianInverseTransposed of the two cells (diagonal)
#pragma unroll
  for (int d = 0; d < DIM; ++d) { jit1[d] = 1.0 / (G.up_i[d] - G.lo_i[d]); jit2[d] = 1.0 / (G.up_o[d] - G.lo_o[d]); }
  for (int q = 0; q < nq; ++q) {
    double pos[3], w; quad_point<DIM - 1>(P.rule, q, pos, w);
    const double* t1 = tab1 + (size_t)q * NL * 4;
    const double* t2 = tab2 + (size_t)q * NL * 4;
    double phi1[NL], phi2[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { phi1[i] = __ldg(t1 + i * 4); phi2[i] = __ldg(t2 + i * 4); }
    const double factor = w * G.integrationElement();
    if (P.op == MDB_OP_PROPFLOW_COUPLING) {
      double u_s = 0.0, u_n = 0.0;
#pragma unroll
      for (int i = 0; i < NL; ++i) u_s += x1[i] * phi1[i];
#pragma unroll
      for (int i = 0; i < NL; ++i) u_n += x2[i] * phi2[i];
      double u_diff = P.intensity * (u_s - u_n);
#pragma unroll
      for (int i = 0; i < NL; ++i) r1[i] += weight * (u_diff * phi1[i] * factor);
#pragma unroll
      for (int i = 0; i < NL; ++i) r2[i] += weight * (-u_diff * phi2[i] * factor);
      continue;
    }
    double gradphi1[NL][DIM], gradphi2[NL][DIM];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradphi1[i][d] = 0.0 + jit1[d] * __ldg(t1 + i * 4 + 1 + d);
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradphi2[i][d] = 0.0 + jit2[d] * __ldg(t2 + i * 4 + 1 + d);
    }
    double gradu1[DIM], gradu2[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { gradu1[d] = 0.0; gradu2[d] = 0.0; }
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu1[d] += x1[i] * gradphi1[i][d];
#pragma unroll
    for (int i = 0; i < NL; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu2[d] += x2[i] * gradphi2[i][d];
    double u1 = 0.0, u2 = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) u1 += x1[i] * phi1[i];
#pragma unroll
    for (int i = 0; i < NL; ++i) u2 += x2[i] * phi2[i];
    double normal1[DIM], normal2[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { normal1[d] = (d == G.axis) ? (G.side ? 1.0 : -1.0) : 0.0; normal2[d] = normal1[d] * -1; }
    const double jump_u1 = u1 - u2;
    const double jump_u2 = u2 - u1;
    double mean_gradu[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { mean_gradu[d] = gradu1[d] + gradu2[d]; mean_gradu[d] *= 0.5; }
    const double theta = 1;
    const double alpha = P.intensity;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      double mgn = 0.0, gpn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) mgn += mean_gradu[d] * normal1[d];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gpn += gradphi1[i][d] * normal1[d];
      double value = 0.0;
      value -= mgn * phi1[i];
      value -= theta * 0.5 * gpn * jump_u1;
      value += alpha / h_F * jump_u1 * phi1[i];
      r1[i] += weight * (factor * value);
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      double mgn = 0.0, gpn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) mgn += mean_gradu[d] * normal2[d];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gpn += gradphi2[i][d] * normal2[d];
      double value = 0.0;
      value -= mgn * phi2[i];
      value -= theta * 0.5 * gpn * jump_u2;
      value += alpha / h_F * jump_u2 * phi2[i];
      r2[i] += weight * (factor * value);
    }
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) k_coupling_residual(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                           const int8_t* __restrict__ fface, int64_t nfaces, double weight,
                                                           const double* __restrict__ tab, const double* __restrict__ x, double* r)
{
  constexpr int NL = 1 << DIM;
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL], dn[NL];
  cell_dofs_q1<DIM>(Ls, G.ijk, ds); cell_dofs_q1<DIM>(Rs, G.nijk, dn);
  double x1[NL], x2[NL], r1[NL], r2[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) { x1[i] = x[ds[i]]; x2[i] = x[dn[i]]; r1[i] = 0.0; r2[i] = 0.0; }
  alpha_coupling<DIM>(P, G, tab, x1, x2, r1, r2, weight);
#pragma unroll
  for (int i = 0; i < NL; ++i) { atomicAdd(&r[ds[i]], r1[i]); atomicAdd(&r[dn[i]], r2[i]); }
}

// Slot table: for every coupling face the offset inside its CSR row of each of the (2 NL)^2 entries the face touches
// (rows: NL self dofs then NL neighbour dofs; columns likewise).  Built once per (matrix, grid operator, coupling).
template <int DIM>
__global__ void k_coupling_slots(MeshDesc m, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface,
                                 int64_t nfaces, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, uint8_t* slots, int* err)
{
  constexpr int NL = 1 << DIM;
  constexpr int NR = 2 * NL;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nfaces * NR * NR) return;
  const int64_t f = t / (NR * NR);
  const int ir = (int)((t / NR) % NR), jc = (int)(t % NR);
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL], dn[NL];
  cell_dofs_q1<DIM>(Ls, G.ijk, ds); cell_dofs_q1<DIM>(Rs, G.nijk, dn);
  int32_t row = 0, col = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    if (ir == i) row = ds[i];
    if (ir == NL + i) row = dn[i];
    if (jc == i) col = ds[i];
    if (jc == NL + i) col = dn[i];
  }
  const int64_t rs = rowptr[row];
  const int64_t pos = find_slot(rowptr, colidx, row, col);
  if (pos >= rowptr[row + 1] || colidx[pos] != col || pos - rs > 255) { atomicExch(err, 1); slots[t] = 0; return; }
  slots[t] = (uint8_t)(pos - rs);
}

// jacobian_coupling, staged variant (the default).  17 residual evaluations per face, no FMA contraction; shared-memory
// traffic and the FP64 pipe bound this kernel, so everything that does not depend on the perturbation is hoisted out of the
// 17 evaluations WITHOUT changing a single rounding of the numbers that reach the matrix:
//   * per (face, quadrature point, side, i) the block first parks in shared memory
//       phi_i, c_i = (theta*0.5)*gpn_i, px_i = x_i*phi_i, pga_i = x_i*gradphi_i[axis]
//     Only the NORMAL component of the gradients is kept: the face normal is +-e_axis, so the reference's dot products with
//     the normal add exact zeros for the tangential components (a + (+-0) == a; the sign of a zero result is normalised by
//     the leading "0.0 +" exactly as in the reference's accumulation from 0.0);
//   * the base-line sums u_s, gradu_s[axis] of each side are formed once per quadrature point, in the reference's order, by
//     a sequential shuffle chain over the 2^d lanes that hold the products of one (face, side);
//   * a perturbation of dof j changes only the j-th product of j's side: the evaluation thread re-sums that side in the
//     reference's order with its own j-th product and takes the other side's base-line sums as they are (same terms, same
//     order: same bits).
// Mapping: NT threads = FPB faces x NE evaluations; evaluation 0 is the base line, 1+j perturbs dof j.
template <int DIM, int NT>
__global__ void __launch_bounds__(NT, 768 / NT) k_coupling_jacobian_staged(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                                  const int8_t* __restrict__ fface, int64_t nfaces, double weight, double epsilon,
                                                                  const double* __restrict__ tab, const double* __restrict__ x,
                                                                  const int64_t* __restrict__ rowptr, const uint8_t* __restrict__ slots, double* val)
{
  constexpr int NL = 1 << DIM;
  constexpr int NR = 2 * NL;
  constexpr int NE = NR + 1;
  constexpr int FPB = NT / NE;
  struct FaceS { double x[NR]; double jit[2]; double nrm[2]; double ie, aoh; int32_t dof[NR]; int axis, side; };   // jit, nrm: the axis component
  __shared__ FaceS sf[FPB];
  // Shared-memory staging, 16-byte vectors (one LDS.128 / STS.128 per access), double-buffered over the quadrature points:
  //   spc  [buf][face][i] = {phi_i, c_i}             perturbation-independent
  //   sq   [buf][face][i] = {px_i, pga_i}            base-line products
  //   sq   [MINE + t]     = the same pair for the perturbed dof of thread t
  //   sbase[buf][face][s] = {u_s, gradu_s[axis]}     base-line sums of side s
  constexpr int MINE = 2 * FPB * NR;
  __shared__ double2 spc[2 * FPB * NR];
  __shared__ double2 sq[MINE + NT];
  __shared__ double2 sbase[2 * FPB * 2];
  __shared__ double sdown[FPB][NR];
  const int t = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * FPB;
  const bool analytic = P.jac_mode == MDB_JAC_ANALYTIC;
  const bool cvcf = P.op == MDB_OP_CVCF_COUPLING;
  // ---- prologue: one thread per face sets up the geometry and gathers the coefficients
  if (t < FPB && f0 + t < nfaces) {
    FaceGeo<DIM> G; G.init(m, fcell[f0 + t], fface[f0 + t]);
    FaceS& F = sf[t];
    int32_t ds[NL], dn[NL];
    cell_dofs_q1<DIM>(Ls, G.ijk, ds); cell_dofs_q1<DIM>(Rs, G.nijk, dn);
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      F.dof[i] = ds[i]; F.dof[NL + i] = dn[i];
      F.x[i] = analytic ? 0.0 : x[ds[i]]; F.x[NL + i] = analytic ? 0.0 : x[dn[i]];
    }
#pragma unroll
    for (int d = 0; d < DIM; ++d)
      if (d == G.axis) {
        F.jit[0] = 1.0 / (G.up_i[d] - G.lo_i[d]); F.jit[1] = 1.0 / (G.up_o[d] - G.lo_o[d]);
        F.nrm[0] = G.side ? 1.0 : -1.0; F.nrm[1] = F.nrm[0] * -1;
      }
    F.ie = G.integrationElement();
    F.aoh = cvcf ? P.intensity / G.cornerDistance01() : 0.0;
    F.axis = G.axis; F.side = G.side;
  }
  __syncthreads();
  const int fl = t / NE, e = t % NE;
  const bool active = fl < FPB && f0 + fl < nfaces;
  const int j = e - 1;                              // perturbed dof in the face's 2 NL numbering, -1 = base line
  double xp = 0.0, delta = 1.0;                     // perturbed coefficient of dof j
  if (active && j >= 0) {
    if (analytic) xp = 1.0;
    else { const double u = sf[fl].x[j]; delta = epsilon * (1.0 + fabs(u)); xp = u + delta; }
  }
  const int jj = j < 0 ? -1 : (j < NL ? j : j - NL);
  const int jside = j < 0 ? -1 : (j < NL ? 0 : 1);
  double r[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) r[i] = 0.0;
  const int nq = ipow(P.rule.m, DIM - 1);
  // phase-1 item of this thread: (face pf, side ps, shape function pi); the NL items of one (face, side) sit in NL adjacent lanes
  const int pf = t / NR, pr = t % NR, ps = pr / NL, pi = pr % NL;
  const bool pact = pf < FPB && f0 + pf < nfaces;
  const int lane0 = (t & 31) - pi;                  // first lane of this (face, side) group
  for (int q = 0; q < nq; ++q) {
    const int bbase = (q & 1) * FPB * NR;
    double px = 0.0, pga = 0.0;
    if (pact) {
      const FaceS& F = sf[pf];
      const int sideval = ps == 0 ? F.side : 1 - F.side;     // geometryInInside: local[axis] = side; InOutside: 1 - side
      const double* tq = tab + ((size_t)((F.axis * 2 + sideval) * nq + q) * NL + pi) * 4;
      const double phi = __ldg(tq);
      const double ga = 0.0 + F.jit[ps] * __ldg(tq + 1 + F.axis);
      const double gpn = 0.0 + ga * F.nrm[ps];
      const double theta = 1;
      const double xi = F.x[pr];
      const int o = bbase + pf * NR + pr;
      px = xi * phi; pga = xi * ga;
      spc[o] = make_double2(phi, theta * 0.5 * gpn);
      sq[o] = make_double2(px, pga);
    }
    {                                                        // base-line sums of the side, in the reference's order (all lanes shuffle)
      double us = 0.0, gs = 0.0;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        us += __shfl_sync(0xffffffffu, px, lane0 + i);
        gs += __shfl_sync(0xffffffffu, pga, lane0 + i);
      }
      if (pact && pi == 0) sbase[(q & 1) * FPB * 2 + pf * 2 + ps] = make_double2(us, gs);
    }
    if (active && j >= 0) {                                  // this thread's perturbed products, in its own slot
      const FaceS& F = sf[fl];
      const int sideval = jside == 0 ? F.side : 1 - F.side;
      const double* tq = tab + ((size_t)((F.axis * 2 + sideval) * nq + q) * NL + jj) * 4;
      const double ga = 0.0 + F.jit[jside] * __ldg(tq + 1 + F.axis);
      sq[MINE + t] = make_double2(xp * __ldg(tq), xp * ga);
    }
    __syncthreads();
    if (active && !(analytic && e == 0)) {
      const FaceS& F = sf[fl];
      const int fb = bbase + fl * NR;
      double pos[3], w; quad_point<DIM - 1>(P.rule, q, pos, w);
      const double factor = w * F.ie;
      const double2 b0 = sbase[(q & 1) * FPB * 2 + fl * 2], b1 = sbase[(q & 1) * FPB * 2 + fl * 2 + 1];
      double u[2] = {b0.x, b1.x};
      double ga[2] = {b0.y, b1.y};
      if (j >= 0) {                                          // re-sum the perturbed side with this thread's j-th product
        double us = 0.0, gs = 0.0;
        const int sb = fb + jside * NL;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          const double2 a = sq[(i == jj) ? MINE + t : sb + i];
          us += a.x; gs += a.y;
        }
        if (jside == 0) { u[0] = us; ga[0] = gs; } else { u[1] = us; ga[1] = gs; }
      }
      if (!cvcf) {
        const double u_diff = P.intensity * (u[0] - u[1]);
#pragma unroll
        for (int i = 0; i < NL; ++i) r[i] += u_diff * spc[fb + i].x * factor;
#pragma unroll
        for (int i = 0; i < NL; ++i) r[NL + i] += -u_diff * spc[fb + NL + i].x * factor;
      } else {
        const double jump[2] = {u[0] - u[1], u[1] - u[0]};
        double mean_ga = ga[0] + ga[1];
        mean_ga *= 0.5;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const double mgn = 0.0 + mean_ga * F.nrm[s];
          const double aj = F.aoh * jump[s];
#pragma unroll
          for (int i = 0; i < NL; ++i) {
            const double2 pc = spc[fb + s * NL + i];
            double value = 0.0;
            value -= mgn * pc.x;
            value -= pc.y * jump[s];
            value += aj * pc.x;
            r[s * NL + i] += factor * value;
          }
        }
      }
    }
  }
  if (active && e == 0) {
#pragma unroll
    for (int i = 0; i < NR; ++i) sdown[fl][i] = r[i];
  }
  __syncthreads();
  if (!active || e == 0) return;
  const uint8_t* sl = slots + (f0 + fl) * (NR * NR) + j;          // column j of the face's (2NL x 2NL) slot table
  const FaceS& F = sf[fl];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const double c = analytic ? r[i] : (r[i] - sdown[fl][i]) / delta;
    atomicAdd(&val[rowptr[F.dof[i]] + sl[i * NR]], weight * c);     // mat_ss / mat_ns (j self), mat_sn / mat_nn (j neighbour)
  }
}

// generic.cu: Qk x Qk couplings with at least one side that is not Q1
void residual_coupling_generic(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, int k1, int k2, const FaceList& fl, double weight,
                               const double* d_x, double* d_r);
void jacobian_coupling_generic(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, int k1, int k2, const FaceList& fl, Matrix& M,
                               std::pair<int, int> key, double weight, double eps, const double* d_x);

static SideMap side_map(const Ctx* c, int sp)
{
  const Child& ch = c->children[c->sps[sp].children[0]];
  SideMap s; for (int d = 0; d < 3; ++d) s.lat_n[d] = ch.lat_n[d];
  s.offset = ch.offset; s.lat2dof = ch.d_lat2dof; s.dg = ch.dg ? 1 : 0;
  return s;
}
static const double* face_tables(Ctx* c, const CouplingParams& P)
{
  if (!c->d_ftab) c->d_ftab = dalloc<double>(MDB_FTAB_DOUBLES);
  const int dim = c->mesh.dim;
  int nq = 1; for (int d = 0; d < dim - 1; ++d) nq *= P.rule.m;
  const int total = 2 * dim * nq * (1 << dim);
  if (dim == 2) MDB_LAUNCH(c, k_face_tables<2>, cdiv(total, 128), 128, 0, P.rule, c->d_ftab);
  else MDB_LAUNCH(c, k_face_tables<3>, cdiv(total, 128), 128, 0, P.rule, c->d_ftab);
  return c->d_ftab;
}
static CouplingParams coupling_params(const CouplingDesc& cp)
{
  CouplingParams P; P.op = cp.op_id; P.jac_mode = cp.jac_mode;
  if (cp.op_id == MDB_OP_CVCF_COUPLING) { P.intensity = cp.cvcf.intensity; P.rule = gauss1d(cp.cvcf.quad_order); }
  else { P.intensity = cp.propflow.intensity; P.rule = gauss1d(4); }     // const int intorder = 4, proportionalflowcoupling.hh:43
  return P;
}

void residual_coupling(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r)
{
  const MeshDesc& m = c->mesh;
  for (size_t k = 0; k < go.cps.size(); ++k) {
    const CouplingDesc& cp = c->cps[go.cps[k]];
    const FaceList& fl = go.faces[k];
    if (fl.n == 0) continue;
    if (cp.enriched()) { residual_enriched(c, cp, fl, weight, d_x, d_r); continue; }      // mortar.cu
    if (cp.op_id == MDB_OP_ND_COUPLING) { residual_nd(c, cp, fl, weight, d_x, d_r); continue; }   // nd.cu
    CouplingParams P = coupling_params(cp);
    SideMap Ls = side_map(c, cp.local_sp), Rs = side_map(c, cp.remote_sp);
    const int k1 = c->children[c->sps[cp.local_sp].children[0]].k, k2 = c->children[c->sps[cp.remote_sp].children[0]].k;
    if (k1 != 1 || k2 != 1 || Ls.dg || Rs.dg) { residual_coupling_generic(c, P, Ls, Rs, k1, k2, fl, weight, d_x, d_r); continue; }
    const double* tab = face_tables(c, P);
    if (m.dim == 2) MDB_LAUNCH(c, k_coupling_residual<2>, cdiv(fl.n, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, tab, d_x, d_r);
    else MDB_LAUNCH(c, k_coupling_residual<3>, cdiv(fl.n, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, tab, d_x, d_r);
  }
}

// `late` selects the couplings of one pass of mdb_jacobian: false = those whose entries all lie in irregular rows (they declare a
// coupling pattern on every DOF of the cells they touch) and may therefore run beside the streamed simple rows; true = the
// Neumann-Dirichlet coupling, which declares no pattern (its mat_ss also hits full-stencil rows one vertex layer off the
// interface) and must run after the volume fill.
void jacobian_coupling(Ctx* c, const GridOp& go, Matrix& M, double weight, const double* d_x, bool late)
{
  const MeshDesc& m = c->mesh;
  for (size_t k = 0; k < go.cps.size(); ++k) {
    const CouplingDesc& cp = c->cps[go.cps[k]];
    const FaceList& fl = go.faces[k];
    if (fl.n == 0) continue;
    if ((cp.op_id == MDB_OP_ND_COUPLING) != late) continue;
    if (cp.enriched()) { jacobian_enriched(c, cp, fl, M, weight, d_x, std::pair<int, int>((int)(&go - &c->gos[0]), (int)k)); continue; }        // mortar.cu
    if (cp.op_id == MDB_OP_ND_COUPLING) { jacobian_nd(c, cp, fl, M, weight); continue; }          // nd.cu
    CouplingParams P = coupling_params(cp);
    SideMap Ls = side_map(c, cp.local_sp), Rs = side_map(c, cp.remote_sp);
    const double eps = 1e-8;                    // NumericalJacobianCoupling() default, couplingutilities.hh:63-65
    const int k1 = c->children[c->sps[cp.local_sp].children[0]].k, k2 = c->children[c->sps[cp.remote_sp].children[0]].k;
    if (k1 != 1 || k2 != 1 || Ls.dg || Rs.dg) {
      jacobian_coupling_generic(c, P, Ls, Rs, k1, k2, fl, M, std::pair<int, int>((int)(&go - &c->gos[0]), (int)k), weight, eps, d_x);
      continue;
    }
    const int NL = 1 << m.dim, NR = 2 * NL;
    // slot table of this (grid operator, coupling) in this matrix, built on first use
    const std::pair<int, int> key((int)(&go - &c->gos[0]), (int)k);
    uint8_t* slots = nullptr;
    auto it = M.slot_tables.find(key);
    if (it == M.slot_tables.end()) {
      slots = dalloc<uint8_t>((size_t)fl.n * NR * NR);
      int* d_err = dalloc<int>(1);
      MDB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
      const int64_t total = fl.n * NR * NR;
      if (m.dim == 2) MDB_LAUNCH(c, k_coupling_slots<2>, cdiv(total, 256), 256, 0, m, Ls, Rs, fl.d_cell, fl.d_face, fl.n, M.d_rowptr, M.d_colidx, slots, d_err);
      else MDB_LAUNCH(c, k_coupling_slots<3>, cdiv(total, 256), 256, 0, m, Ls, Rs, fl.d_cell, fl.d_face, fl.n, M.d_rowptr, M.d_colidx, slots, d_err);
      int err = 0;
      MDB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      dfree(d_err);
      if (err) { cudaFree(slots); throw Error(MDB_EINVAL, "coupling entries missing from the matrix pattern (was the matrix created with this grid operator?)"); }
      M.slot_tables[key] = slots;
    } else slots = it->second;
    if (jacobian_coupling_fd_fast(c, P, Ls, Rs, fl, M, key, slots, weight, eps, d_x)) continue;      // planar interface: fdcoupling.cu
    const double* tab = face_tables(c, P);
    const int fpb = 256 / (2 * NL + 1);
    static const int pad_kb = getenv("MDB_COUPLING_PAD_KB") ? atoi(getenv("MDB_COUPLING_PAD_KB")) : 0;   // occupancy experiment knob
    const size_t pad = (size_t)pad_kb * 1024;
    if (pad > 0) {
      static bool configured = false;
      if (!configured) {
        MDB_CUDA(cudaFuncSetAttribute(k_coupling_jacobian_staged<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
        MDB_CUDA(cudaFuncSetAttribute(k_coupling_jacobian_staged<3, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
        configured = true;
      }
    }
    if (m.dim == 2) MDB_LAUNCH(c, (k_coupling_jacobian_staged<2, 256>), cdiv(fl.n, fpb), 256, pad, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, eps, tab,
                               d_x, M.d_rowptr, slots, M.d_val);
    else MDB_LAUNCH(c, (k_coupling_jacobian_staged<3, 256>), cdiv(fl.n, fpb), 256, pad, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, eps, tab,
                    d_x, M.d_rowptr, slots, M.d_val);
  }
}

// ---- lambda terms of Poisson ---------------------------------------------------------------------------
struct LambdaSpec { int cond_kind; uint32_t mask; mdb_function f; mdb_bctype bc; mdb_function j; Rule1D rule; };

// lambda_volume: r_i -= f(x_q) phi_i(x_q) w_q |det J|  (instationarypoissonoperator.hh:103-117), one thread per cell
template <int DIM, int K>
__global__ void k_lambda_volume(MeshDesc m, const uint8_t* __restrict__ sd, LambdaSpec S, SideMap map, double weight, double t, double* r)
{
  constexpr int NL = cpow(K + 1, DIM);
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells()) return;
  if (!cond_applies(S.cond_kind, S.mask, sd[c])) return;
  int ijk[3]; int64_t rr = c;
  ijk[0] = (int)(rr % m.n[0]); rr /= m.n[0]; ijk[1] = (int)(rr % m.n[1]); ijk[2] = (int)(rr / m.n[1]);
  double lo[3], up[3];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { lo[d] = m.coord(d, ijk[d]); up[d] = m.coord(d, ijk[d] + 1); }
  double detj = up[0] - lo[0];
#pragma unroll
  for (int d = 1; d < DIM; ++d) detj *= (up[d] - lo[d]);
  double rl[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) rl[i] = 0.0;
  const int nq = ipow(S.rule.m, DIM);
  for (int q = 0; q < nq; ++q) {
    double x[3], w; quad_point<DIM>(S.rule, q, x, w);
    double xg[3];
#pragma unroll
    for (int d = 0; d < DIM; ++d) xg[d] = lo[d] + x[d] * (up[d] - lo[d]);
    double y = eval_function(S.f, DIM, xg, t);
    double factor = w * detj;
#pragma unroll
    for (int i = 0; i < NL; ++i) rl[i] += weight * (-y * qk_phi<DIM>(K, i, x) * factor);
  }
  int32_t dofs[NL];
  cell_dofs_qk<DIM, K>(map, ijk, dofs);
#pragma unroll
  for (int i = 0; i < NL; ++i) atomicAdd(&r[dofs[i]], rl[i]);
}

// lambda_boundary on the faces of one boundary plane (axis, side): r_i += j(x_q) phi_i w_q |J_f| where isNeumann
// (instationarypoissonoperator.hh:142-167).  One thread per boundary face.
// All boundary planes of a subproblem go through ONE launch (the kernel is latency bound): job = (axis, side), blocks [block0, ...)
struct PlaneJobs { int n; int axis[6], side[6], block0[7]; };
// one launch for all Poisson subproblems of the grid operator: a boundary face is evaluated for the subproblems that apply to its cell
// (participant order), instead of one launch per subproblem in which the faces of the other subdomains only exit
#define MDB_LAMBDA_JOBS 4
struct LambdaJobs { int n; LambdaSpec S[MDB_LAMBDA_JOBS]; SideMap map[MDB_LAMBDA_JOBS]; };
template <int DIM, int K>
__global__ void k_lambda_boundary(MeshDesc m, const uint8_t* __restrict__ sd, LambdaJobs L, PlaneJobs J, double weight,
                                  double t, double* r)
{
  constexpr int NL = cpow(K + 1, DIM);
  int jb = 0;
  while (jb + 1 < J.n && (int)blockIdx.x >= J.block0[jb + 1]) ++jb;
  const int axis = J.axis[jb], side = J.side[jb];
  int64_t idx = (blockIdx.x - J.block0[jb]) * (int64_t)blockDim.x + threadIdx.x;
  int t0 = (axis == 0) ? 1 : 0, t1 = (axis == 2) ? 1 : 2;
  int64_t nplane = (int64_t)m.n[t0] * ((DIM > 2) ? m.n[t1] : 1);
  if (idx >= nplane) return;
  int ijk[3] = {0, 0, 0};
  ijk[axis] = side ? m.n[axis] - 1 : 0;
  ijk[t0] = (int)(idx % m.n[t0]);
  if (DIM > 2) ijk[t1] = (int)(idx / m.n[t0]);
  int64_t c = ijk[0] + (int64_t)m.n[0] * (ijk[1] + (int64_t)m.n[1] * ijk[2]);
  const uint8_t sdc = sd[c];
  for (int v = 0; v < L.n; ++v) {
  const LambdaSpec& S = L.S[v]; const SideMap& map = L.map[v];
  if (!cond_applies(S.cond_kind, S.mask, sdc)) continue;
  if (S.bc.kind == MDB_BC_ALL_DIRICHLET) continue;
  double lo[3], up[3];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { lo[d] = m.coord(d, ijk[d]); up[d] = m.coord(d, ijk[d] + 1); }
  double fe = 1.0; bool first = true;
#pragma unroll
  for (int d = 0; d < DIM; ++d) if (d != axis) { double e = up[d] - lo[d]; fe = first ? e : fe * e; first = false; }
  double rl[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) rl[i] = 0.0;
  const int nq = ipow(S.rule.m, DIM - 1);
  for (int q = 0; q < nq; ++q) {
    double pos[3], w; quad_point<DIM - 1>(S.rule, q, pos, w);
    double local[3], xg[3], xe[3];
    int tt = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      if (d == axis) { local[d] = (double)side; xg[d] = side ? up[d] : lo[d]; }
      else { local[d] = pos[tt]; xg[d] = lo[d] + pos[tt] * (up[d] - lo[d]); ++tt; }
    }
    if (eval_bctype(S.bc, DIM, xg, true) != 1) continue;
#pragma unroll
    for (int d = 0; d < DIM; ++d) xe[d] = lo[d] + local[d] * (up[d] - lo[d]);
    double y = eval_function(S.j, DIM, xe, t);
    double factor = w * fe;
#pragma unroll
    for (int i = 0; i < NL; ++i) rl[i] += weight * (y * qk_phi<DIM>(K, i, local) * factor);
  }
  int32_t dofs[NL];
  cell_dofs_qk<DIM, K>(map, ijk, dofs);
#pragma unroll
  for (int i = 0; i < NL; ++i) if (rl[i] != 0.0) atomicAdd(&r[dofs[i]], rl[i]);
  }
}

void residual_boundary(Ctx* c, const GridOp& go, double weight, double* d_r)
{
  const MeshDesc& m = c->mesh;
  LambdaJobs L1, L2; L1.n = 0; L2.n = 0;                  // subproblems on Q1 / on Qk (k = 2) children
  auto launch = [&](const LambdaJobs& L, int K) {
    PlaneJobs J; J.n = 0; J.block0[0] = 0;
    for (int axis = 0; axis < m.dim; ++axis)
      for (int side = 0; side < 2; ++side) {
        // only physical boundary planes: a partition cut is a processor boundary (no boundary terms, globalassembler.hh:275-283)
        if (side == 0 && !m.on_global_lower(axis, 0)) continue;
        if (side == 1 && !m.on_global_upper(axis, m.n[axis])) continue;
        int t0 = (axis == 0) ? 1 : 0, t1 = (axis == 2) ? 1 : 2;
        int64_t nplane = (int64_t)m.n[t0] * ((m.dim > 2) ? m.n[t1] : 1);
        J.axis[J.n] = axis; J.side[J.n] = side; J.block0[J.n + 1] = J.block0[J.n] + cdiv(nplane, 128); ++J.n;
      }
    if (J.n == 0) return;
    const int nb = J.block0[J.n];
    if (m.dim == 2 && K == 1) MDB_LAUNCH(c, (k_lambda_boundary<2, 1>), nb, 128, 0, m, c->d_cell_sd, L, J, weight, c->time, d_r);
    else if (m.dim == 3 && K == 1) MDB_LAUNCH(c, (k_lambda_boundary<3, 1>), nb, 128, 0, m, c->d_cell_sd, L, J, weight, c->time, d_r);
    else if (m.dim == 2) MDB_LAUNCH(c, (k_lambda_boundary<2, 2>), nb, 128, 0, m, c->d_cell_sd, L, J, weight, c->time, d_r);
    else MDB_LAUNCH(c, (k_lambda_boundary<3, 2>), nb, 128, 0, m, c->d_cell_sd, L, J, weight, c->time, d_r);
  };
  for (int s : go.sps) {
    const SubProblemDesc& sp = c->sps[s];
    if (sp.op_id != MDB_OP_POISSON) continue;
    LambdaSpec S; S.cond_kind = sp.cond_kind; S.mask = sp.mask; S.f = sp.poisson.f; S.bc = sp.poisson.bc; S.j = sp.poisson.j;
    S.rule = gauss1d(sp.poisson.quad_order);
    SideMap map = side_map(c, s);
    const int K = c->children[sp.children[0]].k;
    if (sp.poisson.f.kind != MDB_FN_ZERO) {       // f == 0 contributes exact zeros
      if (m.dim == 2 && K == 1) MDB_LAUNCH(c, (k_lambda_volume<2, 1>), cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
      else if (m.dim == 3 && K == 1) MDB_LAUNCH(c, (k_lambda_volume<3, 1>), cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
      else if (m.dim == 2) MDB_LAUNCH(c, (k_lambda_volume<2, 2>), cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
      else MDB_LAUNCH(c, (k_lambda_volume<3, 2>), cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
    }
    if (sp.poisson.bc.kind == MDB_BC_ALL_DIRICHLET) continue;   // no Neumann face anywhere
    LambdaJobs& L = (K == 1) ? L1 : L2;
    if (L.n == MDB_LAMBDA_JOBS) { launch(L, K == 1 ? 1 : 2); L.n = 0; }
    L.S[L.n] = S; L.map[L.n] = map; ++L.n;
  }
  if (L1.n > 0) launch(L1, 1);
  if (L2.n > 0) launch(L2, 2);
}

} // namespace mdb
