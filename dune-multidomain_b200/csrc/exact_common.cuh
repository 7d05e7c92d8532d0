// exact_common.cuh — device helpers shared by the translation units that are compiled with --fmad=false (exact.cu, generic.cu):
// the Qk basis, tensor Gauss rules, interface-face geometry, local-to-global maps.
#pragma once
#include "common.cuh"

namespace mdb {

__host__ __device__ constexpr int cpow(int b, int e) { return e == 0 ? 1 : b * cpow(b, e - 1); }

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

struct SideMap { int lat_n[3]; int64_t offset; const int32_t* lat2dof; int dg; };     // dg: lattice stride K + 1 (QkDG child)

// fdcoupling.cu; false = not applicable (the caller runs k_coupling_jacobian_staged)
bool jacobian_coupling_fd_fast(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, const FaceList& fl, Matrix& M, std::pair<int, int> key,
                               const uint8_t* slots, double weight, double eps, const double* d_x);

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

// local index i of a Qk cell = a0 + (K+1) (a1 + (K+1) a2), lattice point = K * cell + a  ([UPSTREAM] QkLocalCoefficients order)
template <int DIM, int K> __device__ __forceinline__ void cell_dofs_qk(const SideMap& S, const int ijk[3], int32_t* dofs)
{
  constexpr int NL = cpow(K + 1, DIM);
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const int ax = i % (K + 1), ay = (i / (K + 1)) % (K + 1), az = (DIM > 2) ? i / ((K + 1) * (K + 1)) : 0;
    const int st = K + S.dg;
    int64_t p = (st * ijk[0] + ax) + (int64_t)S.lat_n[0] * ((st * ijk[1] + ay) + (int64_t)S.lat_n[1] * (st * ijk[2] + az));
    dofs[i] = (int32_t)(S.offset + S.lat2dof[p]);
  }
}

__device__ __forceinline__ int64_t find_slot(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t row, int32_t col)
{
  int64_t lo = rowptr[row], hi = rowptr[row + 1];
  while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (colidx[mid] < col) lo = mid + 1; else hi = mid; }
  return lo;   // the pattern contains every coupling link by construction
}


} // namespace mdb
