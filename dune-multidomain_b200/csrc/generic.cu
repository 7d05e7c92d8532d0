// generic.cu — Qk spaces beyond Q1 (the reference's testpoisson puts Q2 on the right subdomain, test/testpoisson.cc:168-198).
// COMPILED WITH --fmad=false like exact.cu: the finite-difference coupling Jacobian must reproduce the CPU evaluation.
//
// The Q1 kernels of assemble.cu / exact.cu exploit the 3^d vertex stencil (bit masks, constant row stencils, staged FD
// evaluations).  This file holds the same operations for a general Qk lattice (k = 2 instantiated) in their plain form:
//   k_jacobian_rows_qk   owner-computes volume Jacobian: one thread per matrix row sums its incident cells' element-matrix
//                        rows in ascending cell order (jacobianassemblerengine.hh:271-364 scatter order), columns located by
//                        binary search in the row
//   k_residual_rows_qk   alpha_volume of a linear operator applied as element matrix times local coefficients, per row
//   k_coupling_residual_g / k_coupling_jacobian_g   alpha_coupling / NumericalJacobianCoupling for Q(k1) x Q(k2) faces
//                        (test/proportionalflowcoupling.hh:27-86,113-233; couplingutilities.hh:75-135), one thread per
//                        face resp. per (face, evaluation)
#include "exact_common.cuh"

namespace mdb {

struct ChildG {
  int k, nloc, subdomain;
  int lat_n[3];
  int64_t offset, size;
  const int32_t* lat2dof;
  const int32_t* dof2lat;
};
struct VolSpecG {                      // the subproblems of the grid operator that act on this child (participant order)
  int n;
  int cond_kind[MDB_MAX_SP_PER_CHILD];
  uint32_t mask[MDB_MAX_SP_PER_CHILD];
  int slot[MDB_MAX_SP_PER_CHILD];
};

// Visits the cells incident to lattice point p of child ch in ascending cell index; f(cell ijk, subdomain byte, local index of p)
template <int DIM, typename F>
__device__ __forceinline__ void for_incident_cells(const MeshDesc& m, const uint8_t* __restrict__ sd, const ChildG& ch, const int p[3], F f)
{
  const int k = ch.k;
  int lo_c[3], hi_c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d >= DIM) { lo_c[d] = hi_c[d] = 0; continue; }
    const int c1 = p[d] / k;
    hi_c[d] = (c1 < m.n[d]) ? c1 : m.n[d] - 1;
    lo_c[d] = (p[d] % k == 0 && c1 > 0) ? c1 - 1 : c1;
    if (lo_c[d] > hi_c[d]) lo_c[d] = hi_c[d];
  }
  for (int cz = lo_c[2]; cz <= hi_c[2]; ++cz)
    for (int cy = lo_c[1]; cy <= hi_c[1]; ++cy)
      for (int cx = lo_c[0]; cx <= hi_c[0]; ++cx) {
        const uint8_t s = sd[cx + (int64_t)m.n[0] * (cy + (int64_t)m.n[1] * cz)];
        if (ch.subdomain >= 0 && !((s >> ch.subdomain) & 1)) continue;
        const int c[3] = {cx, cy, cz};
        const int a0 = p[0] - k * cx, a1 = (DIM > 1) ? p[1] - k * cy : 0, a2 = (DIM > 2) ? p[2] - k * cz : 0;
        f(c, s, a0 + (k + 1) * (a1 + (k + 1) * a2));
      }
}

template <int DIM>
__device__ __forceinline__ int32_t cell_dof(const ChildG& ch, const int c[3], int lj)
{
  const int k = ch.k;
  const int b0 = lj % (k + 1), b1 = (lj / (k + 1)) % (k + 1), b2 = (DIM > 2) ? lj / ((k + 1) * (k + 1)) : 0;
  const int64_t q = (k * c[0] + b0) + (int64_t)ch.lat_n[0] * ((k * c[1] + b1) + (int64_t)ch.lat_n[1] * (k * c[2] + b2));
  return (int32_t)(ch.offset + ch.lat2dof[q]);
}

template <int DIM>
__global__ void __launch_bounds__(128) k_jacobian_rows_qk(MeshDesc m, const uint8_t* __restrict__ sd, ChildG ch, VolSpecG V, const double* __restrict__ Ae,
                                                          const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, double* __restrict__ val,
                                                          int overwrite)
{
  const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gi >= ch.size) return;
  const int64_t g = ch.offset + gi;
  const int64_t rs = rowptr[g], re = rowptr[g + 1];
  if (overwrite) for (int64_t q = rs; q < re; ++q) val[q] = 0.0;          // the caller's a = 0, fused (this thread owns the row)
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  for_incident_cells<DIM>(m, sd, ch, p, [&](const int* c, uint8_t s, int li) {
    for (int v = 0; v < V.n; ++v) {
      if (!cond_applies(V.cond_kind[v], V.mask[v], s)) continue;
      const double* A = Ae + (size_t)V.slot[v] * MDB_MAXLOC * MDB_MAXLOC + (size_t)li * ch.nloc;
      for (int lj = 0; lj < ch.nloc; ++lj) {
        const int32_t col = cell_dof<DIM>(ch, c, lj);
        int64_t lo = rs, hi = re;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (colidx[mid] < col) lo = mid + 1; else hi = mid; }
        val[lo] += A[lj];                                                     // the pattern holds every volume link by construction
      }
    }
  });
}

template <int DIM>
__global__ void __launch_bounds__(128) k_residual_rows_qk(MeshDesc m, const uint8_t* __restrict__ sd, ChildG ch, VolSpecG V, const double* __restrict__ Ae,
                                                          const double* __restrict__ x, double* __restrict__ res)
{
  const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gi >= ch.size) return;
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  double sum = 0.0;
  for_incident_cells<DIM>(m, sd, ch, p, [&](const int* c, uint8_t s, int li) {
    for (int v = 0; v < V.n; ++v) {
      if (!cond_applies(V.cond_kind[v], V.mask[v], s)) continue;
      const double* A = Ae + (size_t)V.slot[v] * MDB_MAXLOC * MDB_MAXLOC + (size_t)li * ch.nloc;
      double cell = 0.0;
      for (int lj = 0; lj < ch.nloc; ++lj) cell += A[lj] * x[cell_dof<DIM>(ch, c, lj)];
      sum += cell;
    }
  });
  res[ch.offset + gi] += sum;
}

static ChildG child_g(const Ctx* c, int i)
{
  const Child& ch = c->children[i];
  ChildG g; g.k = ch.k; g.nloc = ch.nloc(c->mesh.dim); g.subdomain = ch.subdomain;
  for (int d = 0; d < 3; ++d) g.lat_n[d] = ch.lat_n[d];
  g.offset = ch.offset; g.size = ch.size; g.lat2dof = ch.d_lat2dof; g.dof2lat = ch.d_dof2lat;
  return g;
}
static VolSpecG vol_spec_g(const Ctx* c, const GridOp& go, int child)
{
  VolSpecG V; V.n = 0;
  for (size_t i = 0; i < go.sps.size(); ++i) {
    const SubProblemDesc& sp = c->sps[go.sps[i]];
    if (sp.children[0] != child || !sp.doAlphaVolume()) continue;
    MDB_REQUIRE(V.n < MDB_MAX_SP_PER_CHILD, MDB_EINVAL, "too many subproblems on one child space");
    V.cond_kind[V.n] = sp.cond_kind; V.mask[V.n] = sp.mask; V.slot[V.n] = (int)i; ++V.n;
  }
  return V;
}

// all rows of child `child` (launch_element_matrices must have run)
void jacobian_volume_generic(Ctx* c, const GridOp& go, Matrix& M, int child, bool overwrite)
{
  const MeshDesc& m = c->mesh;
  ChildG ch = child_g(c, child);
  if (ch.size == 0) return;
  VolSpecG V = vol_spec_g(c, go, child);
  if (V.n == 0) {
    if (overwrite) MDB_CUDA(cudaMemsetAsync(M.d_val + M.child_begin[child], 0, (M.child_end[child] - M.child_begin[child]) * sizeof(double), c->stream));
    return;
  }
  if (m.dim == 2) MDB_LAUNCH(c, k_jacobian_rows_qk<2>, cdiv(ch.size, 128), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite);
  else MDB_LAUNCH(c, k_jacobian_rows_qk<3>, cdiv(ch.size, 128), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite);
}

void residual_volume_generic(Ctx* c, const GridOp& go, int child, const double* d_x, double* d_r)
{
  const MeshDesc& m = c->mesh;
  ChildG ch = child_g(c, child);
  VolSpecG V = vol_spec_g(c, go, child);
  if (ch.size == 0 || V.n == 0) return;
  if (m.dim == 2) MDB_LAUNCH(c, k_residual_rows_qk<2>, cdiv(ch.size, 128), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, d_x, d_r);
  else MDB_LAUNCH(c, k_residual_rows_qk<3>, cdiv(ch.size, 128), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, d_x, d_r);
}

// ---- couplings between Q(K1) and Q(K2) ------------------------------------------------------------------------------------
// Face tables as in exact.cu (values and reference gradients of the shape functions at the face quadrature points, per face
// axis and side of the cell), one table per polynomial degree.
#define MDB_FTAB_G_DOUBLES (6 * 25 * 27 * 4)
template <int DIM, int K> __global__ void k_face_tables_g(Rule1D R, double* tab)
{
  constexpr int NL = cpow(K + 1, DIM);
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
  qk_grad<DIM>(K, i, local, js);
  double* o = tab + (size_t)t * 4;
  o[0] = qk_phi<DIM>(K, i, local); o[1] = js[0]; o[2] = js[1]; o[3] = js[2];
}

// alpha_coupling with NL1 shape functions on the inside and NL2 on the outside; same statements, same order as the
// reference (and as exact.cu's Q1 x Q1 version)
template <int DIM, int K1, int K2>
__device__ void alpha_coupling_g(const CouplingParams& P, const FaceGeo<DIM>& G, const double* __restrict__ tabA, const double* __restrict__ tabB,
                                 const double* x1, const double* x2, double* r1, double* r2, double weight)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM);
  const int nq = ipow(P.rule.m, DIM - 1);
  const double h_F = G.cornerDistance01();
  const double* tab1 = tabA + (size_t)((G.axis * 2 + G.side) * nq) * NL1 * 4;          // geometryInInside: local[axis] = side
  const double* tab2 = tabB + (size_t)((G.axis * 2 + (1 - G.side)) * nq) * NL2 * 4;    // geometryInOutside: 1 - side
  double jit1[DIM], jit2[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { jit1[d] = 1.0 / (G.up_i[d] - G.lo_i[d]); jit2[d] = 1.0 / (G.up_o[d] - G.lo_o[d]); }
  for (int q = 0; q < nq; ++q) {
    double pos[3], w; quad_point<DIM - 1>(P.rule, q, pos, w);
    const double* t1 = tab1 + (size_t)q * NL1 * 4;
    const double* t2 = tab2 + (size_t)q * NL2 * 4;
    const double factor = w * G.integrationElement();
    if (P.op == MDB_OP_PROPFLOW_COUPLING) {
      double u_s = 0.0, u_n = 0.0;
      for (int i = 0; i < NL1; ++i) u_s += x1[i] * __ldg(t1 + i * 4);
      for (int i = 0; i < NL2; ++i) u_n += x2[i] * __ldg(t2 + i * 4);
      const double u_diff = P.intensity * (u_s - u_n);
      for (int i = 0; i < NL1; ++i) r1[i] += weight * (u_diff * __ldg(t1 + i * 4) * factor);
      for (int i = 0; i < NL2; ++i) r2[i] += weight * (-u_diff * __ldg(t2 + i * 4) * factor);
      continue;
    }
    double gradu1[DIM], gradu2[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) { gradu1[d] = 0.0; gradu2[d] = 0.0; }
    for (int i = 0; i < NL1; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu1[d] += x1[i] * (0.0 + jit1[d] * __ldg(t1 + i * 4 + 1 + d));
    for (int i = 0; i < NL2; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu2[d] += x2[i] * (0.0 + jit2[d] * __ldg(t2 + i * 4 + 1 + d));
    double u1 = 0.0, u2 = 0.0;
    for (int i = 0; i < NL1; ++i) u1 += x1[i] * __ldg(t1 + i * 4);
    for (int i = 0; i < NL2; ++i) u2 += x2[i] * __ldg(t2 + i * 4);
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
    for (int i = 0; i < NL1; ++i) {
      double mgn = 0.0, gpn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) mgn += mean_gradu[d] * normal1[d];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gpn += (0.0 + jit1[d] * __ldg(t1 + i * 4 + 1 + d)) * normal1[d];
      const double phi = __ldg(t1 + i * 4);
      double value = 0.0;
      value -= mgn * phi;
      value -= theta * 0.5 * gpn * jump_u1;
      value += alpha / h_F * jump_u1 * phi;
      r1[i] += weight * (factor * value);
    }
    for (int i = 0; i < NL2; ++i) {
      double mgn = 0.0, gpn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) mgn += mean_gradu[d] * normal2[d];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gpn += (0.0 + jit2[d] * __ldg(t2 + i * 4 + 1 + d)) * normal2[d];
      const double phi = __ldg(t2 + i * 4);
      double value = 0.0;
      value -= mgn * phi;
      value -= theta * 0.5 * gpn * jump_u2;
      value += alpha / h_F * jump_u2 * phi;
      r2[i] += weight * (factor * value);
    }
  }
}

template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(64) k_coupling_residual_g(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                            const int8_t* __restrict__ fface, int64_t nfaces, double weight, const double* __restrict__ tab1,
                                                            const double* __restrict__ tab2, const double* __restrict__ x, double* r)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM);
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL1], dn[NL2];
  cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds); cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
  double x1[NL1], x2[NL2], r1[NL1], r2[NL2];
  for (int i = 0; i < NL1; ++i) { x1[i] = x[ds[i]]; r1[i] = 0.0; }
  for (int i = 0; i < NL2; ++i) { x2[i] = x[dn[i]]; r2[i] = 0.0; }
  alpha_coupling_g<DIM, K1, K2>(P, G, tab1, tab2, x1, x2, r1, r2, weight);
  for (int i = 0; i < NL1; ++i) atomicAdd(&r[ds[i]], r1[i]);
  for (int i = 0; i < NL2; ++i) atomicAdd(&r[dn[i]], r2[i]);
}

// offsets inside the CSR rows of the (NL1 + NL2)^2 entries a face touches (rows / columns: inside dofs, then outside dofs)
template <int DIM, int K1, int K2>
__global__ void k_coupling_slots_g(MeshDesc m, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces,
                                   const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, uint8_t* slots, int* err)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM), NR = NL1 + NL2;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nfaces * NR * NR) return;
  const int64_t f = t / (NR * NR);
  const int ir = (int)((t / NR) % NR), jc = (int)(t % NR);
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL1], dn[NL2];
  cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds); cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
  int32_t row = 0, col = 0;
  for (int i = 0; i < NL1; ++i) { if (ir == i) row = ds[i]; if (jc == i) col = ds[i]; }
  for (int i = 0; i < NL2; ++i) { if (ir == NL1 + i) row = dn[i]; if (jc == NL1 + i) col = dn[i]; }
  const int64_t rs = rowptr[row];
  const int64_t pos = find_slot(rowptr, colidx, row, col);
  if (pos >= rowptr[row + 1] || colidx[pos] != col || pos - rs > 255) { atomicExch(err, 1); slots[t] = 0; return; }
  slots[t] = (uint8_t)(pos - rs);
}

// jacobian_coupling (couplingutilities.hh:75-135).  One thread per (face, evaluation): evaluation 0 is the base line ("down"),
// evaluation 1+j perturbs inside dof j (j < NL1) or outside dof j - NL1; the base line of a face reaches its perturbation
// threads through shared memory; perturbation thread j owns column j of mat_ss/mat_ns (inside) or mat_sn/mat_nn (outside).
#define CJG_THREADS 128
template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(CJG_THREADS) k_coupling_jacobian_g(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                                     const int8_t* __restrict__ fface, int64_t nfaces, double weight, double epsilon,
                                                                     const double* __restrict__ tab1, const double* __restrict__ tab2,
                                                                     const double* __restrict__ x, const int64_t* __restrict__ rowptr,
                                                                     const uint8_t* __restrict__ slots, double* val)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM), NR = NL1 + NL2;
  constexpr int NE = NR + 1;
  constexpr int FPB = CJG_THREADS / NE;               // faces per block (>= 2 for every instantiated case)
  __shared__ double sdown[FPB][NR];
  const int fl = threadIdx.x / NE, e = threadIdx.x % NE;
  const int64_t f = (int64_t)blockIdx.x * FPB + fl;
  const bool active = fl < FPB && f < nfaces;
  const int j = e - 1;                                // perturbation index, -1 = base line
  const bool analytic = P.jac_mode == MDB_JAC_ANALYTIC;
  int32_t ds[NL1], dn[NL2];
  double r1[NL1], r2[NL2];
  double delta = 1.0;
  if (active) {
    FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
    cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds); cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
    double u1[NL1], u2[NL2];
    for (int i = 0; i < NL1; ++i) { u1[i] = analytic ? 0.0 : x[ds[i]]; r1[i] = 0.0; }
    for (int i = 0; i < NL2; ++i) { u2[i] = analytic ? 0.0 : x[dn[i]]; r2[i] = 0.0; }
    if (analytic) {
      // the shipped coupling residuals are linear in (u_s, u_n): column j = alpha_coupling(e_j), no division
      if (j >= 0 && j < NL1) u1[j] = 1.0;
      if (j >= NL1) u2[j - NL1] = 1.0;
    } else if (j >= 0) {
      if (j < NL1) { delta = epsilon * (1.0 + fabs(u1[j])); u1[j] += delta; }
      else { delta = epsilon * (1.0 + fabs(u2[j - NL1])); u2[j - NL1] += delta; }
    }
    if (!(analytic && e == 0)) alpha_coupling_g<DIM, K1, K2>(P, G, tab1, tab2, u1, u2, r1, r2, 1.0);
    if (e == 0) {
      for (int i = 0; i < NL1; ++i) sdown[fl][i] = r1[i];
      for (int i = 0; i < NL2; ++i) sdown[fl][NL1 + i] = r2[i];
    }
  }
  __syncthreads();
  if (!active || e == 0) return;
  const uint8_t* sl = slots + f * (NR * NR) + j;      // column j of the face's slot table
  for (int i = 0; i < NL1; ++i) {
    const double c = analytic ? r1[i] : (r1[i] - sdown[fl][i]) / delta;
    atomicAdd(&val[rowptr[ds[i]] + sl[i * NR]], weight * c);                      // mat_ss / mat_sn
  }
  for (int i = 0; i < NL2; ++i) {
    const double c = analytic ? r2[i] : (r2[i] - sdown[fl][NL1 + i]) / delta;
    atomicAdd(&val[rowptr[dn[i]] + sl[(NL1 + i) * NR]], weight * c);              // mat_ns / mat_nn
  }
}

static void face_tables_g(Ctx* c, const CouplingParams& P, int k1, int k2, const double** t1, const double** t2)
{
  if (!c->d_ftab_g) c->d_ftab_g = dalloc<double>(2 * MDB_FTAB_G_DOUBLES);
  const int dim = c->mesh.dim;
  int nq = 1; for (int d = 0; d < dim - 1; ++d) nq *= P.rule.m;
  MDB_REQUIRE(nq <= 25, MDB_EINVAL, "coupling quadrature rule too large for the face tables");
  double* tab[2] = {c->d_ftab_g, c->d_ftab_g + MDB_FTAB_G_DOUBLES};
  const int ks[2] = {k1, k2};
  for (int s = 0; s < 2; ++s) {
    int nl = 1; for (int d = 0; d < dim; ++d) nl *= ks[s] + 1;
    const int total = 2 * dim * nq * nl;
    if (dim == 2 && ks[s] == 1) MDB_LAUNCH(c, (k_face_tables_g<2, 1>), cdiv(total, 128), 128, 0, P.rule, tab[s]);
    else if (dim == 2) MDB_LAUNCH(c, (k_face_tables_g<2, 2>), cdiv(total, 128), 128, 0, P.rule, tab[s]);
    else if (ks[s] == 1) MDB_LAUNCH(c, (k_face_tables_g<3, 1>), cdiv(total, 128), 128, 0, P.rule, tab[s]);
    else MDB_LAUNCH(c, (k_face_tables_g<3, 2>), cdiv(total, 128), 128, 0, P.rule, tab[s]);
  }
  *t1 = tab[0]; *t2 = tab[1];
}

// dispatch over (dim, k1, k2) with k1, k2 in {1, 2}
#define MDB_DISPATCH_K(DIMV, K1V, K2V, CALL)                                  \
  do {                                                                        \
    if (DIMV == 2 && K1V == 1 && K2V == 2) { CALL(2, 1, 2); }                 \
    else if (DIMV == 2 && K1V == 2 && K2V == 1) { CALL(2, 2, 1); }            \
    else if (DIMV == 2 && K1V == 2 && K2V == 2) { CALL(2, 2, 2); }            \
    else if (DIMV == 2) { CALL(2, 1, 1); }                                    \
    else if (K1V == 1 && K2V == 2) { CALL(3, 1, 2); }                         \
    else if (K1V == 2 && K2V == 1) { CALL(3, 2, 1); }                         \
    else if (K1V == 2 && K2V == 2) { CALL(3, 2, 2); }                         \
    else { CALL(3, 1, 1); }                                                   \
  } while (0)

void residual_coupling_generic(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, int k1, int k2, const FaceList& fl, double weight,
                               const double* d_x, double* d_r)
{
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "coupling kernels exist for Q1 and Q2 spaces");
  const MeshDesc& m = c->mesh;
  const double *t1, *t2; face_tables_g(c, P, k1, k2, &t1, &t2);
#define CALL(D, A, B) MDB_LAUNCH(c, (k_coupling_residual_g<D, A, B>), cdiv(fl.n, 64), 64, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, t1, t2, d_x, d_r)
  MDB_DISPATCH_K(m.dim, k1, k2, CALL);
#undef CALL
}

void jacobian_coupling_generic(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, int k1, int k2, const FaceList& fl, Matrix& M,
                               std::pair<int, int> key, double weight, double eps, const double* d_x)
{
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "coupling kernels exist for Q1 and Q2 spaces");
  const MeshDesc& m = c->mesh;
  int nl1 = 1, nl2 = 1;
  for (int d = 0; d < m.dim; ++d) { nl1 *= k1 + 1; nl2 *= k2 + 1; }
  const int NR = nl1 + nl2, NE = NR + 1;
  uint8_t* slots = nullptr;
  auto it = M.slot_tables.find(key);
  if (it == M.slot_tables.end()) {
    slots = dalloc<uint8_t>((size_t)fl.n * NR * NR);
    int* d_err = dalloc<int>(1);
    MDB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
    const int64_t total = fl.n * NR * NR;
#define CALL(D, A, B) MDB_LAUNCH(c, (k_coupling_slots_g<D, A, B>), cdiv(total, 256), 256, 0, m, Ls, Rs, fl.d_cell, fl.d_face, fl.n, M.d_rowptr, M.d_colidx, slots, d_err)
    MDB_DISPATCH_K(m.dim, k1, k2, CALL);
#undef CALL
    int err = 0;
    MDB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(d_err);
    if (err) { cudaFree(slots); throw Error(MDB_EINVAL, "coupling entries missing from the matrix pattern (was the matrix created with this grid operator?)"); }
    M.slot_tables[key] = slots;
  } else slots = it->second;
  const double *t1, *t2; face_tables_g(c, P, k1, k2, &t1, &t2);
  const int fpb = CJG_THREADS / NE;
#define CALL(D, A, B) MDB_LAUNCH(c, (k_coupling_jacobian_g<D, A, B>), cdiv(fl.n, fpb), CJG_THREADS, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, eps, t1, t2, d_x, M.d_rowptr, slots, M.d_val)
  MDB_DISPATCH_K(m.dim, k1, k2, CALL);
#undef CALL
}

} // namespace mdb
