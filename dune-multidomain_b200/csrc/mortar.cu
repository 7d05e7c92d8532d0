// mortar.cu — EnrichedCoupling with a coupling (mortar) space: MortarPoissonCoupling of test/testmortarpoisson.cc:117-250.
// COMPILED WITH --fmad=false like exact.cu / generic.cu: the finite-difference Jacobians
// (NumericalJacobianEnrichedCouplingTo{First,Second}SubProblem, couplingutilities.hh:292-484) must reproduce the CPU evaluation.
//
// Per intersection on which the coupling fires (coupling.hh:178-183: local subproblem on the inside cell, remote one on the
// outside cell) the reference calls alpha_enriched_coupling_first(inside, mortar) and _second(outside, mortar)
// (residualassemblerfunctors.hh:219-257); both are alpha_generic<sign> (:156-244):
//   r_i   += w [ -(n.grad u) v_i - (u - lambda) (n.grad v_i) + 2 gamma (u - lambda) v_i ]
//   r_c,i += w [ (n.grad u) - 2 gamma (u - lambda) ] mu_i,        gamma = gamma_0 / |face|,  n = outer normal of the INSIDE cell
// with the element basis evaluated through geometryInInside (sign > 0) or geometryInOutside (sign < 0).
//
//   k_mortar_residual   one thread per intersection: both sides, r_s / r_n / r_c added atomically
//   k_mortar_jacobian   one thread per (intersection, side, perturbed coefficient): base line + perturbed evaluation, column j of
//                       mat_ss|mat_cs (j in the element) or mat_sc|mat_cc (j in the mortar space), couplingutilities.hh:338-386
// This is a correctness configuration ("small"), interface-sized work: plain kernels, columns by binary search.
#include "exact_common.cuh"
#include <algorithm>
#include <cmath>

namespace mdb {

#define MORTAR_MAXCOMP 2
struct MortarSpec {
  double gamma_0; int jac_mode;
  int ncomp;                           // component blocks of a power coupling space (test/testpowermortarpoisson.cc:186-189), <= MORTAR_MAXCOMP
  Rule1D rule1, rule2;                 // qorder = 2 max(order(lfsu), order(lfsu_c)) per side, testmortarpoisson.cc:193-195
  int lkind, rkind; uint32_t lmask, rmask;    // the coupling space's own predicate (gfs.contains(is))
  const double* tab;                   // shape-function tables per (sign, face orientation), k_mortar_tables
};

// corner i of face (axis, side) of cell ijk: bit t of i = coordinate along the t-th axis other than `axis` ([UPSTREAM] Yasp)
template <int DIM> __device__ __forceinline__ void face_dofs(const SideMap& C, const int ijk[3], int axis, int side, int32_t* dofs)
{
  constexpr int NC = 1 << (DIM - 1);
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    int v[3] = {ijk[0], ijk[1], ijk[2]}; int t = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) { if (d == axis) v[d] += side; else { v[d] += (i >> t) & 1; ++t; } }
    dofs[i] = (int32_t)(C.offset + C.lat2dof[v[0] + (int64_t)C.lat_n[0] * (v[1] + (int64_t)C.lat_n[1] * v[2])]);
  }
}

// alpha_generic<sign> (testmortarpoisson.cc:156-244): same statements, same order as the CPU restatement.  Until session 4 of round 2 every
// thread evaluated the Qk shape functions and gradients at its quadrature points itself (qk_phi / qk_grad: Lagrange products with
// divisions, local-memory arrays): BASELINE configs[3] at 256^3 spent 6.2 ms in the Jacobian of the 65 536 interface faces.
// The functions below are __noinline__ on purpose: inlined into mortar_column (two calls around a dynamically indexed store into u / uc)
// nvcc 12.9 produced NaN coefficients for the second call whenever NL > 4 (observed on B200, both Jacobian modes; the host build of the
// same source under ASan/UBSan is clean).  As a real call the function is correct for every instantiation.
// ---- the same operator from tables (session 4 of round 2) -------------------------------------------------------------------------------
// What alpha_generic evaluates at a quadrature point — the element's shape functions and reference gradients at the face point seen from
// the inside or the outside cell, the mortar shape functions, the weight — depends on the face only through its orientation (axis, side):
// on a structured mesh these are at most 2 DIM x 2 small tables.  They are computed ONCE per coupling by the same device functions
// (qk_phi, qk_grad, quad_point: the same bits), and the Jacobian / residual threads then run alpha_generic's statements in the same
// order on table values.  Everything that depends on the cell (1 / extent per axis, gamma, the integration element) is still evaluated per
// thread exactly as before, so every number that reaches the matrix keeps its rounding sequence (the parity tests hold at 1e-12 against the
// oracle's finite differences, as with the function above).  BASELINE configs[3] at 256^3: Jacobian 6.2 ms with per-thread shape functions.
#define MT_NQ 9
#define MT_NL 27
#define MT_V 0
#define MT_JS (MT_NQ * MT_NL)
#define MT_MU (MT_JS + MT_NQ * MT_NL * 3)
#define MT_W (MT_MU + MT_NQ * 4)
#define MT_STRIDE 1024
template <int DIM>
__global__ void k_mortar_tables(Rule1D rule1, Rule1D rule2, int k1, int k2, double* tab)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * 6 * MT_NQ * MT_NL) return;
  const int i = t % MT_NL, q = (t / MT_NL) % MT_NQ, o = (t / (MT_NL * MT_NQ)) % 6, sgn = t / (MT_NL * MT_NQ * 6);
  const int K = sgn ? k2 : k1;
  const Rule1D& R = sgn ? rule2 : rule1;
  const int nq = ipow(R.m, DIM - 1);
  int NL = 1; for (int d = 0; d < DIM; ++d) NL *= K + 1;
  if (o >= 2 * DIM || q >= nq) return;
  const int axis = o / 2, side = o % 2;
  double pos[3] = {0.0, 0.0, 0.0}, w; quad_point<DIM - 1>(R, q, pos, w);
  double element_pos[3]; int tt = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) element_pos[d] = (d == axis) ? (double)(sgn ? 1 - side : side) : pos[tt++];     // FaceGeo::localInside / localOutside
  double* T = tab + (size_t)(sgn * 6 + o) * MT_STRIDE;
  if (i < NL) {
    T[MT_V + q * MT_NL + i] = qk_phi<DIM>(K, i, element_pos);
    double js[3] = {0.0, 0.0, 0.0}; qk_grad<DIM>(K, i, element_pos, js);
    for (int d = 0; d < 3; ++d) T[MT_JS + (q * MT_NL + i) * 3 + d] = js[d];
  }
  if (i < (1 << (DIM - 1))) T[MT_MU + q * 4 + i] = qk_phi<DIM - 1>(1, i, pos);
  if (i == 0) T[MT_W + q] = w;
}
template <int DIM, int K>
__device__ __noinline__ void mortar_alpha_tab(int sign, double gamma_0, int nq, const double* __restrict__ tab, const FaceGeo<DIM>& G, const double* x,
                                              const double* x_c, int nc, double* r, double* r_c, double weight)
{
  constexpr int NL = cpow(K + 1, DIM), NCB = 1 << (DIM - 1);
  const double* __restrict__ T = tab + (size_t)((sign > 0 ? 0 : 6) + 2 * G.axis + G.side) * MT_STRIDE;
  const double gamma = gamma_0 / G.integrationElement();
  double jit[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) jit[d] = sign > 0 ? 1.0 / (G.up_i[d] - G.lo_i[d]) : 1.0 / (G.up_o[d] - G.lo_o[d]);
  double normal[DIM];
#pragma unroll
  for (int d = 0; d < DIM; ++d) normal[d] = (d == G.axis) ? (G.side ? 1.0 : -1.0) : 0.0;
  if constexpr (NL <= 9) {
    // small local spaces (Q1; Q2 in 2-D): coefficients and accumulators live in registers for the whole call — the caller's arrays are
    // stack memory (this function is a real call), read once and written once instead of once per quadrature point.  Every r[i] still
    // receives its terms one quadrature point after the other.
    double xr[NL], xcr[NCB], ra[NL], rca[NCB];
#pragma unroll
    for (int i = 0; i < NL; ++i) { xr[i] = x[i]; ra[i] = r[i]; }
#pragma unroll
    for (int i = 0; i < NCB; ++i) { xcr[i] = i < nc ? x_c[i] : 0.0; rca[i] = i < nc ? r_c[i] : 0.0; }
    for (int q = 0; q < nq; ++q) {
      const double w = T[MT_W + q];
      double u = 0.0, gradu[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu[d] = 0.0;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        u += xr[i] * T[MT_V + q * MT_NL + i];
#pragma unroll
        for (int d = 0; d < DIM; ++d) gradu[d] += xr[i] * (0.0 + jit[d] * T[MT_JS + (q * MT_NL + i) * 3 + d]);
      }
      double lambda = 0.0;
#pragma unroll
      for (int i = 0; i < NCB; ++i) if (i < nc) lambda += xcr[i] * T[MT_MU + q * 4 + i];
      const double factor = w * G.integrationElement();
      double ngu = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) ngu += normal[d] * gradu[d];
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const double vi = T[MT_V + q * MT_NL + i];
        double ngv = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) ngv += normal[d] * (0.0 + jit[d] * T[MT_JS + (q * MT_NL + i) * 3 + d]);
        ra[i] += weight * ((-ngu * vi - (u - lambda) * ngv + 2 * gamma * (u - lambda) * vi) * factor);
      }
#pragma unroll
      for (int i = 0; i < NCB; ++i) if (i < nc) rca[i] += weight * ((ngu - 2 * gamma * (u - lambda)) * T[MT_MU + q * 4 + i] * factor);
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) r[i] = ra[i];
#pragma unroll
    for (int i = 0; i < NCB; ++i) if (i < nc) r_c[i] = rca[i];
  } else {
  for (int q = 0; q < nq; ++q) {
    const double w = T[MT_W + q];
    double u = 0.0, gradu[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gradu[d] = 0.0;
    for (int i = 0; i < NL; ++i) {
      u += x[i] * T[MT_V + q * MT_NL + i];
#pragma unroll
      for (int d = 0; d < DIM; ++d) gradu[d] += x[i] * (0.0 + jit[d] * T[MT_JS + (q * MT_NL + i) * 3 + d]);
    }
    double lambda = 0.0;
    for (int i = 0; i < nc; ++i) lambda += x_c[i] * T[MT_MU + q * 4 + i];
    const double factor = w * G.integrationElement();
    double ngu = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) ngu += normal[d] * gradu[d];
    for (int i = 0; i < NL; ++i) {
      const double vi = T[MT_V + q * MT_NL + i];
      double ngv = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) ngv += normal[d] * (0.0 + jit[d] * T[MT_JS + (q * MT_NL + i) * 3 + d]);
      r[i] += weight * ((-ngu * vi - (u - lambda) * ngv + 2 * gamma * (u - lambda) * vi) * factor);
    }
    for (int i = 0; i < nc; ++i) r_c[i] += weight * ((ngu - 2 * gamma * (u - lambda)) * T[MT_MU + q * 4 + i] * factor);
  }
  }
}

// the operator on a power coupling space: the body of alpha_generic once per component block, each with its own multiplier
// coefficients x_c[b * NC ..] and its own rows r_c[b * NC ..]; the element residual collects every block's terms
template <int DIM, int K>
__device__ __forceinline__ void mortar_alpha_power(int sign, const MortarSpec& P, const Rule1D& R, const FaceGeo<DIM>& G, const double* x, const double* x_c,
                                                   int nc, double* r, double* r_c, double weight)
{
  constexpr int NC = 1 << (DIM - 1);
  const int nq = ipow(R.m, DIM - 1);
  for (int b = 0; b < P.ncomp; ++b) mortar_alpha_tab<DIM, K>(sign, P.gamma_0, nq, P.tab, G, x, x_c + b * NC, nc, r, r_c + b * NC, weight);
}

struct MortarMaps { SideMap c[MORTAR_MAXCOMP]; };

template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(64) k_mortar_residual(MeshDesc m, const uint8_t* __restrict__ sd, MortarSpec P, SideMap Ls, SideMap Rs, MortarMaps Cs,
                                                        const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces, double weight,
                                                        const double* __restrict__ x, double* r)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM), NC = 1 << (DIM - 1), NCT = MORTAR_MAXCOMP * NC;
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL1], dn[NL2], dc[NCT];
  cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds); cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
  const uint8_t s_in = sd[fcell[f]], s_out = sd[G.nijk[0] + (int64_t)m.n[0] * (G.nijk[1] + (int64_t)m.n[1] * G.nijk[2])];
  const int nc = (cond_applies(P.lkind, P.lmask, s_in) && cond_applies(P.rkind, P.rmask, s_out)) ? NC : 0;   // lfs_c.bind(is): empty if !contains(is)
  if (nc) for (int b = 0; b < P.ncomp; ++b) face_dofs<DIM>(Cs.c[b], G.ijk, G.axis, G.side, dc + b * NC);
  double x1[NL1], x2[NL2], xc[NCT], r1[NL1], r2[NL2], rc[NCT];
  for (int i = 0; i < NL1; ++i) { x1[i] = x[ds[i]]; r1[i] = 0.0; }
  for (int i = 0; i < NL2; ++i) { x2[i] = x[dn[i]]; r2[i] = 0.0; }
  for (int i = 0; i < NCT; ++i) { xc[i] = (nc && i < P.ncomp * NC) ? x[dc[i]] : 0.0; rc[i] = 0.0; }
  mortar_alpha_power<DIM, K1>(+1, P, P.rule1, G, x1, xc, nc, r1, rc, weight);
  mortar_alpha_power<DIM, K2>(-1, P, P.rule2, G, x2, xc, nc, r2, rc, weight);
  for (int i = 0; i < NL1; ++i) atomicAdd(&r[ds[i]], r1[i]);
  for (int i = 0; i < NL2; ++i) atomicAdd(&r[dn[i]], r2[i]);
  if (nc) for (int i = 0; i < P.ncomp * NC; ++i) atomicAdd(&r[dc[i]], rc[i]);
}

// one column of jacobian_enriched_coupling_{first,second} (couplingutilities.hh:305-388 / :404-484)
template <int DIM, int K>
__device__ void mortar_column(int sign, const MortarSpec& P, const Rule1D& R, const FaceGeo<DIM>& G, const int32_t* de, const int32_t* dc, int nc, int j,
                              double weight, double epsilon, const double* __restrict__ x, const int64_t* __restrict__ rowptr,
                              const int32_t* __restrict__ colidx, double* val, uint16_t* slots, int64_t sbase, int64_t sstride, int have_slots)
{
  constexpr int NL = cpow(K + 1, DIM), NC = MORTAR_MAXCOMP * (1 << (DIM - 1));
  const int nct = nc * P.ncomp;                                                                        // coefficients of the (power) coupling space on this face
  const bool analytic = P.jac_mode == MDB_JAC_ANALYTIC;
  double u[NL], uc[NC], down[NL], down_c[NC], up[NL], up_c[NC];
  for (int i = 0; i < NL; ++i) { u[i] = analytic ? 0.0 : x[de[i]]; down[i] = 0.0; up[i] = 0.0; }
  for (int i = 0; i < NC; ++i) { uc[i] = (analytic || i >= nct) ? 0.0 : x[dc[i]]; down_c[i] = 0.0; up_c[i] = 0.0; }
  double delta = 1.0;
  if (!analytic) mortar_alpha_power<DIM, K>(sign, P, R, G, u, uc, nc, down, down_c, 1.0);            // base line
  if (j < NL) {                                                                                        // jiggle in self
    if (!analytic) delta = epsilon * (1.0 + fabs(u[j]));
    u[j] += delta;
  } else {                                                                                             // jiggle in coupling
    if (!analytic) delta = epsilon * (1.0 + fabs(uc[j - NL]));
    uc[j - NL] += delta;
  }
  mortar_alpha_power<DIM, K>(sign, P, R, G, u, uc, nc, up, up_c, 1.0);
  const int32_t col = (j < NL) ? de[j] : dc[j - NL];
  // where entry (row, col) lives inside its row: searched once per (matrix, coupling) and kept as a 16-bit offset per (thread, entry),
  // entry-major so that the threads of a warp read consecutive offsets (the binary searches were most of this kernel's time)
  auto slot = [&](int i, int32_t row) -> int64_t {
    const int64_t rs = rowptr[row];
    if (have_slots) return rs + slots[sbase + (int64_t)i * sstride];
    const int64_t pos = find_slot(rowptr, colidx, row, col);
    if (slots) slots[sbase + (int64_t)i * sstride] = (uint16_t)(pos - rs);
    return pos;
  };
  for (int i = 0; i < NL; ++i) {                                                                      // mat_ss / mat_sc
    const double cval = (up[i] - down[i]) / delta;
    atomicAdd(&val[slot(i, de[i])], weight * cval);
  }
  for (int i = 0; i < nct; ++i) {                                                                     // mat_cs / mat_cc
    const double cval = (up_c[i] - down_c[i]) / delta;
    atomicAdd(&val[slot(NL + i, dc[i])], weight * cval);
  }
}

template <int DIM, int K1, int K2>
__global__ void __launch_bounds__(64) k_mortar_jacobian(MeshDesc m, const uint8_t* __restrict__ sd, MortarSpec P, SideMap Ls, SideMap Rs, MortarMaps Cs,
                                                        const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces, double weight,
                                                        double epsilon, const double* __restrict__ x, const int64_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx, double* val, uint16_t* slots, int have_slots)
{
  constexpr int NL1 = cpow(K1 + 1, DIM), NL2 = cpow(K2 + 1, DIM), NCB = 1 << (DIM - 1);
  const int NC = P.ncomp * NCB;                        // coefficients of the (power) coupling space per face
  const int NP = NL1 + NC + NL2 + NC;                  // perturbations per intersection: first (element, mortar), second (element, mortar)
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nfaces * NP) return;
  // perturbation-major: the threads of a warp perturb the SAME coefficient on 32 consecutive faces — one side, one code path, uniform
  // table loads (face-major, 24 different perturbations per warp: ncu counted 15.9 active threads per warp instruction and twice the
  // instructions, 1.02 ms alone at 256^3)
  const int64_t f = t % nfaces; const int e = (int)(t / nfaces);
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  const uint8_t s_in = sd[fcell[f]], s_out = sd[G.nijk[0] + (int64_t)m.n[0] * (G.nijk[1] + (int64_t)m.n[1] * G.nijk[2])];
  const int nc = (cond_applies(P.lkind, P.lmask, s_in) && cond_applies(P.rkind, P.rmask, s_out)) ? NCB : 0;
  int32_t dc[MORTAR_MAXCOMP * NCB];
  if (nc) for (int b = 0; b < P.ncomp; ++b) face_dofs<DIM>(Cs.c[b], G.ijk, G.axis, G.side, dc + b * NCB);
  if (e < NL1 + NC) {
    if (e >= NL1 && nc == 0) return;
    int32_t ds[NL1]; cell_dofs_qk<DIM, K1>(Ls, G.ijk, ds);
    mortar_column<DIM, K1>(+1, P, P.rule1, G, ds, dc, nc, e, weight, epsilon, x, rowptr, colidx, val, slots, t, nfaces * NP, have_slots);
  } else {
    const int j = e - (NL1 + NC);
    if (j >= NL2 && nc == 0) return;
    int32_t dn[NL2]; cell_dofs_qk<DIM, K2>(Rs, G.nijk, dn);
    mortar_column<DIM, K2>(-1, P, P.rule2, G, dn, dc, nc, j, weight, epsilon, x, rowptr, colidx, val, slots, t, nfaces * NP, have_slots);
  }
}

static SideMap side_map_child(const Ctx* c, int child)
{
  const Child& ch = c->children[child];
  SideMap s; for (int d = 0; d < 3; ++d) s.lat_n[d] = ch.lat_n[d];
  s.offset = ch.offset; s.lat2dof = ch.d_lat2dof; s.dg = ch.dg ? 1 : 0;
  return s;
}

static MortarSpec mortar_spec(Ctx* c, const CouplingDesc& cp, int k1, int k2)
{
  MortarSpec P; P.gamma_0 = cp.mortar.gamma_0; P.jac_mode = cp.jac_mode;
  const Child& cc = c->children[cp.coupling_child];
  P.rule1 = gauss1d(2 * (k1 > cc.k ? k1 : cc.k)); P.rule2 = gauss1d(2 * (k2 > cc.k ? k2 : cc.k));
  P.lkind = cc.lkind; P.lmask = cc.lmask; P.rkind = cc.rkind; P.rmask = cc.rmask;
  P.ncomp = cp.ncomp;
  // tables of this coupling: built by the first call, on the context stream, and complete before any kernel that reads them is queued
  P.tab = nullptr;
  const int nq1 = (int)std::lround(std::pow((double)P.rule1.m, c->mesh.dim - 1)), nq2 = (int)std::lround(std::pow((double)P.rule2.m, c->mesh.dim - 1));
  MDB_REQUIRE(nq1 <= MT_NQ && nq2 <= MT_NQ, MDB_EINVAL, "enriched coupling: quadrature rule larger than the shape-function tables");
  {
    const int id = (int)(&cp - c->cps.data());
    auto it = c->mortar_tabs.find(id);
    if (it == c->mortar_tabs.end()) {
      double* d = dalloc<double>((size_t)12 * MT_STRIDE);
      MDB_CUDA(cudaMemsetAsync(d, 0, (size_t)12 * MT_STRIDE * sizeof(double), c->stream));
      const int total = 2 * 6 * MT_NQ * MT_NL;
      if (c->mesh.dim == 2) MDB_LAUNCH(c, k_mortar_tables<2>, cdiv(total, 128), 128, 0, P.rule1, P.rule2, k1, k2, d);
      else MDB_LAUNCH(c, k_mortar_tables<3>, cdiv(total, 128), 128, 0, P.rule1, P.rule2, k1, k2, d);
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      c->mortar_tabs[id] = d;
      P.tab = d;
    } else P.tab = it->second;
  }
  return P;
}

static MortarMaps mortar_maps(const Ctx* c, const CouplingDesc& cp)
{
  MDB_REQUIRE(cp.ncomp >= 1 && cp.ncomp <= MORTAR_MAXCOMP, MDB_EINVAL, "power coupling spaces: at most 2 component blocks");
  MortarMaps M;
  for (int b = 0; b < MORTAR_MAXCOMP; ++b) M.c[b] = side_map_child(c, cp.coupling_child + (b < cp.ncomp ? b : 0));
  return M;
}

#define MDB_DISPATCH_MORTAR(DIMV, K1V, K2V, CALL)                             \
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

void residual_enriched(Ctx* c, const CouplingDesc& cp, const FaceList& fl, double weight, const double* d_x, double* d_r)
{
  const MeshDesc& m = c->mesh;
  MDB_REQUIRE(m.dim >= 2, MDB_EINVAL, "enriched couplings need dim >= 2");
  const int cl = c->sps[cp.local_sp].children[0], cr = c->sps[cp.remote_sp].children[0];
  const int k1 = c->children[cl].k, k2 = c->children[cr].k;
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "enriched coupling kernels exist for Q1 and Q2 spaces");
  const MortarSpec P = mortar_spec(c, cp, k1, k2);
  const SideMap Ls = side_map_child(c, cl), Rs = side_map_child(c, cr);
  const MortarMaps Cs = mortar_maps(c, cp);
#define CALL(D, A, B) MDB_LAUNCH(c, (k_mortar_residual<D, A, B>), cdiv(fl.n, 64), 64, 0, m, c->d_cell_sd, P, Ls, Rs, Cs, fl.d_cell, fl.d_face, fl.n, weight, d_x, d_r)
  MDB_DISPATCH_MORTAR(m.dim, k1, k2, CALL);
#undef CALL
}

void jacobian_enriched(Ctx* c, const CouplingDesc& cp, const FaceList& fl, Matrix& M, double weight, const double* d_x, std::pair<int, int> key)
{
  const MeshDesc& m = c->mesh;
  MDB_REQUIRE(m.dim >= 2, MDB_EINVAL, "enriched couplings need dim >= 2");
  const int cl = c->sps[cp.local_sp].children[0], cr = c->sps[cp.remote_sp].children[0];
  const int k1 = c->children[cl].k, k2 = c->children[cr].k;
  MDB_REQUIRE(k1 >= 1 && k1 <= 2 && k2 >= 1 && k2 <= 2, MDB_EINVAL, "enriched coupling kernels exist for Q1 and Q2 spaces");
  const MortarSpec P = mortar_spec(c, cp, k1, k2);
  const SideMap Ls = side_map_child(c, cl), Rs = side_map_child(c, cr);
  const MortarMaps Cs = mortar_maps(c, cp);
  int nl1 = 1, nl2 = 1;
  for (int d = 0; d < m.dim; ++d) { nl1 *= k1 + 1; nl2 *= k2 + 1; }
  const int64_t total = fl.n * (nl1 + nl2 + 2 * cp.ncomp * (1 << (m.dim - 1)));
  const double eps = 1e-8;                      // NumericalJacobianEnrichedCouplingTo*SubProblem() default, couplingutilities.hh:296-298
  // slot table of this (grid operator, coupling) in this matrix: filled by the first call (which searches), read by the later ones
  static const bool no_slots = getenv("MDB_MORTAR_SEARCH") != nullptr;                 // A/B switch: binary search per entry on every call
  const int smax = std::max(nl1, nl2) + cp.ncomp * (1 << (m.dim - 1));
  uint16_t* slots = nullptr; int have = 0;
  if (!no_slots && M.max_row < 65536) {
    auto it = M.slot_tables.find(key);
    if (it == M.slot_tables.end()) { slots = dalloc<uint16_t>((size_t)total * smax); M.slot_tables[key] = reinterpret_cast<uint8_t*>(slots); }
    else { slots = reinterpret_cast<uint16_t*>(it->second); have = 1; }
  }
#define CALL(D, A, B) MDB_LAUNCH(c, (k_mortar_jacobian<D, A, B>), cdiv(total, 64), 64, 0, m, c->d_cell_sd, P, Ls, Rs, Cs, fl.d_cell, fl.d_face, fl.n, weight, eps, d_x, M.d_rowptr, M.d_colidx, M.d_val, slots, have)
  MDB_DISPATCH_MORTAR(m.dim, k1, k2, CALL);
#undef CALL
}

} // namespace mdb
