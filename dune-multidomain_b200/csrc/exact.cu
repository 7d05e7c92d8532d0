// exact.cu — kernels whose arithmetic follows the reference's operation order term by term.
// COMPILED WITH --fmad=false (see Makefile): no FMA contraction, so that the finite-difference coupling Jacobians
// (NumericalJacobianCoupling, couplingutilities.hh:59-141, eps = 1e-8) reproduce the CPU evaluation; a one-ulp change in a
// residual evaluation is amplified to ~1e-8 in (up-down)/delta.
//
//   k_element_matrices  local operator tables: jacobian_volume of Poisson ([UPSTREAM] Laplace::jacobian_volume) and of
//                       L2 / test/simpletimeoperator.hh:68-101 on the (congruent) cells of a structured mesh
//   k_coupling_residual alpha_coupling of ContinuousValueContinuousFlowCoupling (test/proportionalflowcoupling.hh:113-233)
//                       and ProportionalFlowCoupling (:27-86), one thread per interface face
//   k_coupling_jacobian jacobian_coupling: one warp per interface face, lane 0 = base line, lane 1+j = perturbation j
//   k_lambda_volume / k_lambda_boundary   source and Neumann terms of Poisson (test/instationarypoissonoperator.hh:103-167)
#include "common.cuh"

namespace mdb {

// ---- Qk basis ([UPSTREAM] QkLocalBasis) ---------------------------------------------------------------
__device__ __forceinline__ double qk_p(int k, int i, double x)
{
  double result = 1.0;
  for (int j = 0; j <= k; ++j) if (j != i) result *= (k * x - j) / (i - j);
  return result;
}
__device__ __forceinline__ double qk_dp(int k, int i, double x)
{
  double result = 0.0;
  for (int j = 0; j <= k; ++j) if (j != i) {
    double prod = (k * 1.0) / (i - j);
    for (int l = 0; l <= k; ++l) if (l != i && l != j) prod *= (k * x - l) / (i - l);
    result += prod;
  }
  return result;
}
template <int DIM> __device__ __forceinline__ void qk_alpha(int k, int i, int a[3])
{
  for (int d = 0; d < 3; ++d) { a[d] = (d < DIM) ? i % (k + 1) : 0; i /= (k + 1); }
}
template <int DIM> __device__ __forceinline__ double qk_phi(int k, int i, const double* x)
{
  int a[3]; qk_alpha<DIM>(k, i, a);
  double out = 1.0;
#pragma unroll
  for (int j = 0; j < DIM; ++j) out *= qk_p(k, a[j], x[j]);
  return out;
}
template <int DIM> __device__ __forceinline__ void qk_grad(int k, int i, const double* x, double* out)
{
  int a[3]; qk_alpha<DIM>(k, i, a);
#pragma unroll
  for (int j = 0; j < DIM; ++j) {
    out[j] = qk_dp(k, a[j], x[j]);
#pragma unroll
    for (int l = 0; l < DIM; ++l) if (l != j) out[j] *= qk_p(k, a[l], x[l]);
  }
}

template <int DIM> __device__ __forceinline__ void quad_point(const Rule1D& R, int idx, double* x, double& w)
{
  // tensor rule, axis 0 fastest; weight = ((w0*w1)*w2)
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    int i = idx % R.m; idx /= R.m;
    x[d] = R.x[i];
    w = (d == 0) ? R.w[i] : w * R.w[i];
  }
  if (DIM == 0) w = 1.0;
}
__device__ __forceinline__ int ipow(int b, int e) { int r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

// ---- element matrices --------------------------------------------------------------------------------
struct AeSpec { int n; int op[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN]; int k[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN];
                Rule1D rule[MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN]; };

// grid: one block per slot; thread t = i*nloc + j computes entry (i,j) with the reference's summation order over the
// quadrature points.  ext[d] = up-lo of the (congruent) cells.
template <int DIM> __global__ void k_element_matrices(AeSpec S, double ext0, double ext1, double ext2, double weight, double* Ae)
{
  const int slot = blockIdx.x;
  const int k = S.k[slot];
  const int nloc = ipow(k + 1, DIM);
  const int t = threadIdx.x;
  if (t >= nloc * nloc) return;
  const int i = t / nloc, j = t % nloc;
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
  for (int q = 0; q < nq; ++q) {
    double x[3], w; quad_point<DIM>(R, q, x, w);
    double factor = w * detj;
    if (S.op[slot] == MDB_OP_POISSON) {
      double gi[3], gj[3];
      qk_grad<DIM>(k, i, x, gi); qk_grad<DIM>(k, j, x, gj);
      double dot = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) { double a_i = 0.0 + jit[d] * gi[d]; double a_j = 0.0 + jit[d] * gj[d]; dot += a_j * a_i; }
      a += weight * (dot * factor);
    } else {   // MDB_OP_L2
      double pi = qk_phi<DIM>(k, i, x), pj = qk_phi<DIM>(k, j, x);
      a += weight * (pj * pi * factor);
    }
  }
  Ae[slot * MDB_MAXLOC * MDB_MAXLOC + i * nloc + j] = a;
}

void launch_element_matrices(Ctx* c, const std::vector<int>& sps, double weight)
{
  const MeshDesc& m = c->mesh;
  AeSpec S; S.n = (int)sps.size();
  MDB_REQUIRE(S.n <= MDB_MAX_SP_PER_CHILD * MDB_MAX_CHILDREN, MDB_EINVAL, "too many subproblems in one grid operator");
  if (S.n == 0) return;
  for (int i = 0; i < S.n; ++i) {
    const SubProblemDesc& sp = c->sps[sps[i]];
    S.op[i] = sp.op_id; S.k[i] = c->children[sp.children[0]].k; S.rule[i] = gauss1d(sp.quad_order());
  }
  double ext[3] = {1, 1, 1};
  for (int d = 0; d < m.dim; ++d) ext[d] = m.coord(d, 1) - m.coord(d, 0);
  if (m.dim == 2) MDB_LAUNCH(c, k_element_matrices<2>, S.n, 32 * 32, 0, S, ext[0], ext[1], ext[2], weight, c->d_Ae);
  else MDB_LAUNCH(c, k_element_matrices<3>, S.n, 32 * 32, 0, S, ext[0], ext[1], ext[2], weight, c->d_Ae);
}

// ---- interface geometry ------------------------------------------------------------------------------
template <int DIM> struct FaceGeo {
  int ijk[3], nijk[3], axis, side;
  double lo_i[3], up_i[3], lo_o[3], up_o[3];
  __device__ void init(const MeshDesc& m, int64_t cell, int face)
  {
    int64_t c = cell;
    ijk[0] = (int)(c % m.n[0]); c /= m.n[0]; ijk[1] = (int)(c % m.n[1]); ijk[2] = (int)(c / m.n[1]);
    axis = face / 2; side = face % 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) nijk[d] = ijk[d];
    nijk[axis] += side ? 1 : -1;
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      lo_i[d] = m.coord(d, ijk[d]); up_i[d] = m.coord(d, ijk[d] + 1);
      lo_o[d] = m.coord(d, nijk[d]); up_o[d] = m.coord(d, nijk[d] + 1);
    }
  }
  __device__ double integrationElement() const
  {
    double v = 1.0; bool first = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d) if (d != axis) { double e = up_i[d] - lo_i[d]; v = first ? e : v * e; first = false; }
    return v;
  }
  __device__ double cornerDistance01() const
  {
    int t0 = (axis == 0) ? 1 : 0;
    double d = lo_i[t0] - up_i[t0];
    return sqrt(d * d);
  }
  __device__ void localInside(const double* pos, double* local) const
  {
    int t = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) local[d] = (d == axis) ? (double)side : pos[t++];
  }
  __device__ void localOutside(const double* pos, double* local) const
  {
    int t = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) local[d] = (d == axis) ? (double)(1 - side) : pos[t++];
  }
};

struct CouplingParams {
  int op; int jac_mode;
  double intensity;
  Rule1D rule;
};

// alpha_coupling restated for Q1 on both sides (NL = 2^DIM shape functions per side); r_s, r_n accumulate weight*value.
template <int DIM>
__device__ void alpha_coupling(const CouplingParams& P, const FaceGeo<DIM>& G, const double* x1, const double* x2,
                               double* r1, double* r2, double weight)
{
  constexpr int NL = 1 << DIM;
  const int nq = ipow(P.rule.m, DIM - 1);
  const double h_F = G.cornerDistance01();
  for (int q = 0; q < nq; ++q) {
    double pos[3], w; quad_point<DIM - 1>(P.rule, q, pos, w);
    double local1[3], local2[3];
    G.localInside(pos, local1); G.localOutside(pos, local2);
    double phi1[NL], phi2[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { phi1[i] = qk_phi<DIM>(1, i, local1); phi2[i] = qk_phi<DIM>(1, i, local2); }
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
      double js[3];
      qk_grad<DIM>(1, i, local1, js);
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradphi1[i][d] = 0.0 + (1.0 / (G.up_i[d] - G.lo_i[d])) * js[d];
      qk_grad<DIM>(1, i, local2, js);
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradphi2[i][d] = 0.0 + (1.0 / (G.up_o[d] - G.lo_o[d])) * js[d];
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

struct SideMap { int lat_n[3]; int64_t offset; const int32_t* lat2dof; };

template <int DIM> __device__ __forceinline__ void cell_dofs_q1(const SideMap& S, const int ijk[3], int32_t* dofs)
{
  constexpr int NL = 1 << DIM;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    int ax = i & 1, ay = (i >> 1) & 1, az = (DIM > 2) ? (i >> 2) & 1 : 0;
    int64_t p = (ijk[0] + ax) + (int64_t)S.lat_n[0] * ((ijk[1] + ay) + (int64_t)S.lat_n[1] * (ijk[2] + az));
    dofs[i] = (int32_t)(S.offset + S.lat2dof[p]);
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) k_coupling_residual(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                           const int8_t* __restrict__ fface, int64_t nfaces, double weight,
                                                           const double* __restrict__ x, double* r)
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
  alpha_coupling<DIM>(P, G, x1, x2, r1, r2, weight);
#pragma unroll
  for (int i = 0; i < NL; ++i) { atomicAdd(&r[ds[i]], r1[i]); atomicAdd(&r[dn[i]], r2[i]); }
}

__device__ __forceinline__ int64_t find_slot(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t row, int32_t col)
{
  int64_t lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (colidx[mid] < col) lo = mid + 1; else hi = mid; }
  return lo;   // the pattern contains every coupling link by construction
}

// One warp per interface face.  lane 0: base line ("down"); lane 1+j, j<NL: perturb self dof j; lane 1+NL+j: perturb
// neighbour dof j.  Each perturbation lane owns one column of mat_ss/mat_ns (self) or mat_sn/mat_nn (neighbour).
template <int DIM>
__global__ void __launch_bounds__(128) k_coupling_jacobian(MeshDesc m, CouplingParams P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                           const int8_t* __restrict__ fface, int64_t nfaces, double weight, double epsilon,
                                                           const double* __restrict__ x, const int64_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ colidx, double* val)
{
  constexpr int NL = 1 << DIM;
  const int lane = threadIdx.x & 31;
  int64_t f = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (f >= nfaces) return;                      // whole warp exits together
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL], dn[NL];
  cell_dofs_q1<DIM>(Ls, G.ijk, ds); cell_dofs_q1<DIM>(Rs, G.nijk, dn);
  double u1[NL], u2[NL], r1[NL], r2[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) { u1[i] = x[ds[i]]; u2[i] = x[dn[i]]; r1[i] = 0.0; r2[i] = 0.0; }
  const int j = lane - 1;                       // perturbation index, -1 = base line
  const bool self = j >= 0 && j < NL, nb = j >= NL && j < 2 * NL;
  double delta = 1.0;
  if (P.jac_mode == MDB_JAC_ANALYTIC) {
    // the shipped coupling residuals are linear in (u_s,u_n): column j = alpha_coupling(e_j), no division
#pragma unroll
    for (int i = 0; i < NL; ++i) { u1[i] = (self && i == j) ? 1.0 : 0.0; u2[i] = (nb && i == j - NL) ? 1.0 : 0.0; }
  } else {
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (self && i == j) { delta = epsilon * (1.0 + fabs(u1[i])); u1[i] += delta; }
      if (nb && i == j - NL) { delta = epsilon * (1.0 + fabs(u2[i])); u2[i] += delta; }
    }
  }
  alpha_coupling<DIM>(P, G, u1, u2, r1, r2, 1.0);
  double c1[NL], c2[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    double d1 = __shfl_sync(0xffffffffu, r1[i], 0), d2 = __shfl_sync(0xffffffffu, r2[i], 0);
    if (P.jac_mode == MDB_JAC_ANALYTIC) { c1[i] = r1[i]; c2[i] = r2[i]; }
    else { c1[i] = (r1[i] - d1) / delta; c2[i] = (r2[i] - d2) / delta; }
  }
  if (!(self || nb)) return;
  const int32_t col = self ? ds[j] : dn[j - NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    atomicAdd(&val[find_slot(rowptr, colidx, ds[i], col)], weight * c1[i]);     // mat_ss / mat_sn
    atomicAdd(&val[find_slot(rowptr, colidx, dn[i], col)], weight * c2[i]);     // mat_ns / mat_nn
  }
}

static SideMap side_map(const Ctx* c, int sp)
{
  const Child& ch = c->children[c->sps[sp].children[0]];
  SideMap s; for (int d = 0; d < 3; ++d) s.lat_n[d] = ch.lat_n[d];
  s.offset = ch.offset; s.lat2dof = ch.d_lat2dof;
  return s;
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
    CouplingParams P = coupling_params(cp);
    SideMap Ls = side_map(c, cp.local_sp), Rs = side_map(c, cp.remote_sp);
    if (m.dim == 2) MDB_LAUNCH(c, k_coupling_residual<2>, cdiv(fl.n, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, d_x, d_r);
    else MDB_LAUNCH(c, k_coupling_residual<3>, cdiv(fl.n, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, d_x, d_r);
  }
}

void jacobian_coupling(Ctx* c, const GridOp& go, Matrix& M, double weight, const double* d_x)
{
  const MeshDesc& m = c->mesh;
  for (size_t k = 0; k < go.cps.size(); ++k) {
    const CouplingDesc& cp = c->cps[go.cps[k]];
    const FaceList& fl = go.faces[k];
    if (fl.n == 0) continue;
    CouplingParams P = coupling_params(cp);
    SideMap Ls = side_map(c, cp.local_sp), Rs = side_map(c, cp.remote_sp);
    const double eps = 1e-8;                    // NumericalJacobianCoupling() default, couplingutilities.hh:63-65
    int64_t threads = fl.n * 32;
    if (m.dim == 2) MDB_LAUNCH(c, k_coupling_jacobian<2>, cdiv(threads, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, eps, d_x,
                               M.d_rowptr, M.d_colidx, M.d_val);
    else MDB_LAUNCH(c, k_coupling_jacobian<3>, cdiv(threads, 128), 128, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, eps, d_x,
                    M.d_rowptr, M.d_colidx, M.d_val);
  }
}

// ---- lambda terms of Poisson ---------------------------------------------------------------------------
struct LambdaSpec { int cond_kind; uint32_t mask; mdb_function f; mdb_bctype bc; mdb_function j; Rule1D rule; };

// lambda_volume: r_i -= f(x_q) phi_i(x_q) w_q |det J|  (instationarypoissonoperator.hh:103-117), one thread per cell
template <int DIM>
__global__ void k_lambda_volume(MeshDesc m, const uint8_t* __restrict__ sd, LambdaSpec S, SideMap map, double weight, double t, double* r)
{
  constexpr int NL = 1 << DIM;
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
    for (int i = 0; i < NL; ++i) rl[i] += weight * (-y * qk_phi<DIM>(1, i, x) * factor);
  }
  int32_t dofs[NL];
  cell_dofs_q1<DIM>(map, ijk, dofs);
#pragma unroll
  for (int i = 0; i < NL; ++i) atomicAdd(&r[dofs[i]], rl[i]);
}

// lambda_boundary on the faces of one boundary plane (axis, side): r_i += j(x_q) phi_i w_q |J_f| where isNeumann
// (instationarypoissonoperator.hh:142-167).  One thread per boundary face.
template <int DIM>
__global__ void k_lambda_boundary(MeshDesc m, const uint8_t* __restrict__ sd, LambdaSpec S, SideMap map, int axis, int side, double weight,
                                  double t, double* r)
{
  constexpr int NL = 1 << DIM;
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int t0 = (axis == 0) ? 1 : 0, t1 = (axis == 2) ? 1 : 2;
  int64_t nplane = (int64_t)m.n[t0] * ((DIM > 2) ? m.n[t1] : 1);
  if (idx >= nplane) return;
  int ijk[3] = {0, 0, 0};
  ijk[axis] = side ? m.n[axis] - 1 : 0;
  ijk[t0] = (int)(idx % m.n[t0]);
  if (DIM > 2) ijk[t1] = (int)(idx / m.n[t0]);
  int64_t c = ijk[0] + (int64_t)m.n[0] * (ijk[1] + (int64_t)m.n[1] * ijk[2]);
  if (!cond_applies(S.cond_kind, S.mask, sd[c])) return;
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
    for (int i = 0; i < NL; ++i) rl[i] += weight * (y * qk_phi<DIM>(1, i, local) * factor);
  }
  int32_t dofs[NL];
  cell_dofs_q1<DIM>(map, ijk, dofs);
#pragma unroll
  for (int i = 0; i < NL; ++i) if (rl[i] != 0.0) atomicAdd(&r[dofs[i]], rl[i]);
}

void residual_boundary(Ctx* c, const GridOp& go, double weight, double* d_r)
{
  const MeshDesc& m = c->mesh;
  for (int s : go.sps) {
    const SubProblemDesc& sp = c->sps[s];
    if (sp.op_id != MDB_OP_POISSON) continue;
    LambdaSpec S; S.cond_kind = sp.cond_kind; S.mask = sp.mask; S.f = sp.poisson.f; S.bc = sp.poisson.bc; S.j = sp.poisson.j;
    S.rule = gauss1d(sp.poisson.quad_order);
    SideMap map = side_map(c, s);
    if (sp.poisson.f.kind != MDB_FN_ZERO) {       // f == 0 contributes exact zeros
      if (m.dim == 2) MDB_LAUNCH(c, k_lambda_volume<2>, cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
      else MDB_LAUNCH(c, k_lambda_volume<3>, cdiv(m.ncells(), 128), 128, 0, m, c->d_cell_sd, S, map, weight, c->time, d_r);
    }
    if (sp.poisson.bc.kind == MDB_BC_ALL_DIRICHLET) continue;   // no Neumann face anywhere
    for (int axis = 0; axis < m.dim; ++axis)
      for (int side = 0; side < 2; ++side) {
        // only physical boundary planes: a partition cut is a processor boundary (no boundary terms, globalassembler.hh:275-283)
        if (side == 0 && !m.on_global_lower(axis, 0)) continue;
        if (side == 1 && !m.on_global_upper(axis, m.n[axis])) continue;
        int t0 = (axis == 0) ? 1 : 0, t1 = (axis == 2) ? 1 : 2;
        int64_t nplane = (int64_t)m.n[t0] * ((m.dim > 2) ? m.n[t1] : 1);
        if (m.dim == 2) MDB_LAUNCH(c, k_lambda_boundary<2>, cdiv(nplane, 128), 128, 0, m, c->d_cell_sd, S, map, axis, side, weight, c->time, d_r);
        else MDB_LAUNCH(c, k_lambda_boundary<3>, cdiv(nplane, 128), 128, 0, m, c->d_cell_sd, S, map, axis, side, weight, c->time, d_r);
      }
  }
}

} // namespace mdb
