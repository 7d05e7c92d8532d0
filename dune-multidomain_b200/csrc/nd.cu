// nd.cu — ConvectionDiffusionDGNeumannDirichletCoupling<CouplingParameters> (test/neumanndirichletcoupling.hh:41-441, parameter
// class test/testpoisson-dirichlet-neumann.cc:36-69: A = I, b = 0, c = 0), the coupling operator of the Dirichlet-Neumann
// iteration (BASELINE configs[0]).  Compiled with --fmad=false like the other operator files, same statements and order as
// the CPU restatement.
//
// The operator is ONE-DIRECTIONAL: alpha_coupling adds to r_s only (:85-283; r_r is a const reference) and its own analytic
// jacobian_coupling to mat_ss only (:291-432) — the remote coefficients enter the residual as data (explicit coupling).  It
// declares no pattern (:51-52): on conforming spaces every (i, j) of mat_ss is a volume link of the inside cell.
//   k_nd_residual   one thread per intersection the coupling fires on
//   k_nd_jacobian   one thread per (intersection, column j): the four terms of :421-429 summed over the rule, added atomically
#include "exact_common.cuh"

namespace mdb {

struct NDSpec { int mode, weights; double theta, alpha; Rule1D rule; };

template <int DIM> struct NDSetup { double An_F_s[DIM], n_F[DIM], penalty_factor; };

template <int DIM, int K1, int K2>
__device__ __forceinline__ NDSetup<DIM> nd_setup(const NDSpec& P, const FaceGeo<DIM>& G)
{
  NDSetup<DIM> S;
  double vol_s = G.up_i[0] - G.lo_i[0], vol_n = G.up_o[0] - G.lo_o[0];
#pragma unroll
  for (int d = 1; d < DIM; ++d) { vol_s *= (G.up_i[d] - G.lo_i[d]); vol_n *= (G.up_o[d] - G.lo_o[d]); }
  const double h_F = fmin(vol_s, vol_n) / G.integrationElement();          // :128
  double An_F_n[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) S.n_F[d] = (d == G.axis) ? (G.side ? 1.0 : -1.0) : 0.0;
#pragma unroll
  for (int i = 0; i < DIM; ++i) {                                           // A.mv(n_F, An_F) with A = I
    S.An_F_s[i] = 0.0; An_F_n[i] = 0.0;
#pragma unroll
    for (int j = 0; j < DIM; ++j) { S.An_F_s[i] += ((i == j) ? 1.0 : 0.0) * S.n_F[j]; An_F_n[i] += ((i == j) ? 1.0 : 0.0) * S.n_F[j]; }
  }
  double harmonic_average = 1.0;
  if (P.weights == MDB_DG_WEIGHTS_ON) {                                     // :149-155
    double delta_s = 0.0, delta_n = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { delta_s += S.An_F_s[d] * S.n_F[d]; delta_n += An_F_n[d] * S.n_F[d]; }
    harmonic_average = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
  }
  constexpr int degree = (K1 > K2) ? K1 : K2;
  S.penalty_factor = (P.alpha / h_F) * harmonic_average * degree * (degree + DIM - 1);       // :168
  return S;
}

template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(64) k_nd_residual(MeshDesc m, NDSpec P, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell,
                                                    const int8_t* __restrict__ fface, int64_t nfaces, double weight, const double* __restrict__ x, double* r)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM);
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL1], dn[NL2];
  cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds); cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
  double x_s[NL1], x_r[NL2], r_s[NL1];
  for (int i = 0; i < NL1; ++i) { x_s[i] = x[ds[i]]; r_s[i] = 0.0; }
  for (int i = 0; i < NL2; ++i) x_r[i] = x[dn[i]];
  const NDSetup<DIM> S = nd_setup<DIM, K1, K2>(P, G);
  const int nq = ipow(P.rule.m, DIM - 1);
  double jit_s[DIM], jit_n[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) { jit_s[d] = 1.0 / (G.up_i[d] - G.lo_i[d]); jit_n[d] = 1.0 / (G.up_o[d] - G.lo_o[d]); }
  for (int q = 0; q < nq; ++q) {
    double pos[3] = {0.0, 0.0, 0.0}, w; quad_point<DIM - 1>(P.rule, q, pos, w);
    double iplocal_s[3], iplocal_n[3];
    G.localInside(pos, iplocal_s); G.localOutside(pos, iplocal_n);
    double psi_s[NL1];
    for (int i = 0; i < NL1; ++i) psi_s[i] = qk_phi<DIM>(K1, i, iplocal_s);
    const double factor = w * G.integrationElement();
    double normalflux = 0.0;                                                // b = 0
#pragma unroll
    for (int d = 0; d < DIM; ++d) normalflux += 0.0 * S.n_F[d];
    double u_r = 0.0;
    for (int i = 0; i < NL2; ++i) u_r += x_r[i] * qk_phi<DIM>(K2, i, iplocal_n);
    if (P.mode == MDB_ND_MODE_NEUMANN || P.mode == MDB_ND_MODE_BOTH) {      // :194-217
      double gradu_r[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu_r[d] = 0.0;
      for (int i = 0; i < NL2; ++i) {
        double js[3]; qk_grad<DIM>(K2, i, iplocal_n, js);
#pragma unroll
        for (int d = 0; d < DIM; ++d) gradu_r[d] += x_r[i] * (0.0 + jit_n[d] * js[d]);
      }
      double gn = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) gn += gradu_r[d] * S.n_F[d];
      const double j = normalflux * u_r - gn;
      for (int i = 0; i < NL1; ++i) r_s[i] += weight * (j * psi_s[i] * factor);
      if (P.mode == MDB_ND_MODE_NEUMANN) continue;
    }
    double u_s = 0.0;
    for (int i = 0; i < NL1; ++i) u_s += x_s[i] * psi_s[i];
    double tgrad_s[NL1][DIM], gradu_s[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gradu_s[d] = 0.0;
    for (int i = 0; i < NL1; ++i) {
      double js[3]; qk_grad<DIM>(K1, i, iplocal_s, js);
#pragma unroll
      for (int d = 0; d < DIM; ++d) tgrad_s[i][d] = 0.0 + jit_s[d] * js[d];
    }
    for (int i = 0; i < NL1; ++i)
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu_s[d] += x_s[i] * tgrad_s[i][d];
    const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0, omegaup_r = normalflux >= 0.0 ? 0.0 : 1.0;
    const double term1 = (omegaup_s * u_s + omegaup_r * u_r) * normalflux * factor;
    for (int i = 0; i < NL1; ++i) r_s[i] += weight * (term1 * psi_s[i]);
    double Agu = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) Agu += S.An_F_s[d] * gradu_s[d];
    const double term2 = -factor * Agu;
    for (int i = 0; i < NL1; ++i) r_s[i] += weight * (term2 * psi_s[i]);
    const double term3 = (u_s - u_r) * factor;
    for (int i = 0; i < NL1; ++i) {
      double Agp = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) Agp += S.An_F_s[d] * tgrad_s[i][d];
      r_s[i] += weight * (term3 * P.theta * Agp);
    }
    const double term4 = S.penalty_factor * (u_s - u_r) * factor;
    for (int i = 0; i < NL1; ++i) r_s[i] += weight * (term4 * psi_s[i]);
  }
  for (int i = 0; i < NL1; ++i) atomicAdd(&r[ds[i]], r_s[i]);
}

template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(64) k_nd_jacobian(MeshDesc m, NDSpec P, SideMap Ls, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface,
                                                    int64_t nfaces, double weight, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                    double* val)
{
  constexpr int NL1 = cpow(K1 + 1, DIM);
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nfaces * NL1) return;
  const int64_t f = t / NL1; const int j = (int)(t % NL1);
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL1];
  cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds);
  const NDSetup<DIM> S = nd_setup<DIM, K1, K2>(P, G);
  const int nq = ipow(P.rule.m, DIM - 1);
  double jit_s[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) jit_s[d] = 1.0 / (G.up_i[d] - G.lo_i[d]);
  double col[NL1];
  for (int i = 0; i < NL1; ++i) col[i] = 0.0;
  for (int q = 0; q < nq; ++q) {
    double pos[3] = {0.0, 0.0, 0.0}, w; quad_point<DIM - 1>(P.rule, q, pos, w);
    double iplocal_s[3];
    G.localInside(pos, iplocal_s);
    double normalflux = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) normalflux += 0.0 * S.n_F[d];
    const double omegaup_s = normalflux >= 0.0 ? 1.0 : 0.0;
    const double factor = w * G.integrationElement();
    const double ipfactor = S.penalty_factor * factor;
    const double phi_j = qk_phi<DIM>(K1, j, iplocal_s);
    double js[3]; qk_grad<DIM>(K1, j, iplocal_s, js);
    double Agj = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) Agj += S.An_F_s[d] * (0.0 + jit_s[d] * js[d]);
    const double temp1 = -Agj * factor;
    for (int i = 0; i < NL1; ++i) {                                         // :421-429: I convection, II diffusion, III consistency, IV ip
      const double phi_i = qk_phi<DIM>(K1, i, iplocal_s);
      double ji[3]; qk_grad<DIM>(K1, i, iplocal_s, ji);
      double Agi = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) Agi += S.An_F_s[d] * (0.0 + jit_s[d] * ji[d]);
      col[i] += weight * (omegaup_s * phi_j * normalflux * factor * phi_i);
      col[i] += weight * (temp1 * phi_i);
      col[i] += weight * (phi_j * factor * P.theta * Agi);
      col[i] += weight * (phi_j * ipfactor * phi_i);
    }
  }
  for (int i = 0; i < NL1; ++i) atomicAdd(&val[find_slot(rowptr, colidx, ds[i], ds[j])], col[i]);
}

static SideMap side_map_child(const Ctx* c, int child)
{
  const Child& ch = c->children[child];
  SideMap s; for (int d = 0; d < 3; ++d) s.lat_n[d] = ch.lat_n[d];
  s.offset = ch.offset; s.lat2dof = ch.d_lat2dof; s.dg = ch.dg ? 1 : 0;
  return s;
}

static NDSpec nd_spec(const CouplingDesc& cp, int k1, int k2)
{
  NDSpec P; P.mode = cp.nd.mode; P.weights = cp.nd.weights; P.alpha = cp.nd.alpha;
  P.theta = (cp.nd.method == MDB_DG_METHOD_SIPG) ? -1.0 : 1.0;             // :72-73
  P.rule = gauss1d(cp.nd.intorderadd + 2 * (k1 > k2 ? k1 : k2));           // :107
  return P;
}

#define MDB_DISPATCH_ND(DIMV, K1V, K2V, CALL)                                 \
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

void residual_nd(Ctx* c, const CouplingDesc& cp, const FaceList& fl, double weight, const double* d_x, double* d_r)
{
  const MeshDesc& m = c->mesh;
  MDB_REQUIRE(m.dim >= 2, MDB_EINVAL, "the Neumann-Dirichlet coupling needs dim >= 2");
  const int cl = c->sps[cp.local_sp].children[0], cr = c->sps[cp.remote_sp].children[0];
  const int k1 = c->children[cl].k, k2 = c->children[cr].k;
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "coupling kernels exist for Q1 and Q2 spaces");
  const NDSpec P = nd_spec(cp, k1, k2);
  const SideMap Ls = side_map_child(c, cl), Rs = side_map_child(c, cr);
#define CALL(D, A, B) MDB_LAUNCH(c, (k_nd_residual<D, A, B>), cdiv(fl.n, 64), 64, 0, m, P, Ls, Rs, fl.d_cell, fl.d_face, fl.n, weight, d_x, d_r)
  MDB_DISPATCH_ND(m.dim, k1, k2, CALL);
#undef CALL
}

void jacobian_nd(Ctx* c, const CouplingDesc& cp, const FaceList& fl, Matrix& M, double weight)
{
  if (cp.nd.mode == MDB_ND_MODE_NEUMANN) return;                           // :299-300
  const MeshDesc& m = c->mesh;
  const int cl = c->sps[cp.local_sp].children[0], cr = c->sps[cp.remote_sp].children[0];
  const int k1 = c->children[cl].k, k2 = c->children[cr].k;
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "coupling kernels exist for Q1 and Q2 spaces");
  const NDSpec P = nd_spec(cp, k1, k2);
  const SideMap Ls = side_map_child(c, cl);
  int nl1 = 1; for (int d = 0; d < m.dim; ++d) nl1 *= k1 + 1;
  const int64_t total = fl.n * nl1;
#define CALL(D, A, B) MDB_LAUNCH(c, (k_nd_jacobian<D, A, B>), cdiv(total, 64), 64, 0, m, P, Ls, fl.d_cell, fl.d_face, fl.n, weight, M.d_rowptr, M.d_colidx, M.d_val)
  MDB_DISPATCH_ND(m.dim, k1, k2, CALL);
#undef CALL
}

} // namespace mdb
