// fdcoupling.cu — the finite-difference coupling Jacobian on a PLANAR interface of Q1 cells, register/constant-bank variant.
// COMPILED WITH --fmad=false (see Makefile): every number that reaches the matrix has the bits of the reference's evaluation
// order (NumericalJacobianCoupling, couplingutilities.hh:75-135 around ContinuousValueContinuousFlowCoupling::alpha_coupling,
// test/proportionalflowcoupling.hh:113-233).
//
// Why a second kernel.  k_coupling_jacobian_staged (exact.cu) handles any face list; ncu showed it issue-bound on shared-memory
// traffic and index arithmetic (1.98e8 warp instructions for 65 536 faces, FP64 pipe a quarter of them).  On a planar interface
// — every face has the same (axis, side) and lies in the same lattice plane, the case of every shipped driver — the shape-function
// values at the face quadrature points are the SAME numbers for every face.  They are evaluated once (k_fd_tables, with the very
// qk_phi / qk_grad arithmetic the other kernels use), copied to the host when the plan is built and handed to the kernel as a
// __grid_constant__ parameter block, so that inside the fully unrolled quadrature loop they are constant-bank operands of the
// DMUL / DADD instructions: no load instruction, no index arithmetic.
//
// What is hoisted, without changing one rounding of a number that reaches the matrix:
//   * a node that does not lie on the face has phi = +-0 there, so its products with phi are exact zeros and are skipped
//     (s + (+-0) == s bit for bit; a running sum that starts at +0.0 is never -0.0);
//   * the normal derivative of the shape function of the node OPPOSITE to an on-face node is the exact negative of the
//     on-face node's, so c_off * jump == -(c_on * jump) and the product is formed once (checked bitwise when the plan is built:
//     a plan whose tables violate either identity is refused and the staged kernel runs instead);
//   * the face normal is +-e_axis: the dot products with the normal keep the axis component only (see exact.cu).
// Mapping: a block owns FPB faces; thread = (face, perturbed coefficient j); a warp holds the j of ONE side of 32/NL faces, so
// the side of j — which decides which nodes are on the face — is warp-uniform and compiled in (fd_body<..., JS>).  The
// perturbation-independent products x_i phi_i, x_i dphi_i and the base-line sums of each side are parked in shared memory once
// per block (phase 1); a thread re-sums only the perturbed side in the reference's order with its own j-th product (the address
// of term i is chosen once per thread), evaluates the 2 NL rows for its column, and the base line of ROW j — so that the 17th
// evaluation of the reference's loop costs 1/NR of an evaluation per thread.
#include "exact_common.cuh"
#include <algorithm>
#include <type_traits>

namespace mdb {

#define FD_NT 128                       // threads per block
#define FD_MAXNQ 16

// tables of one plan, device layout: tab[(q*NR + s*NL + i)*4 + {phi, ga, c, -}] ; scal = {jit_s, jit_n}
template <int DIM>
__global__ void k_fd_tables(MeshDesc m, Rule1D R, int32_t cell0, int face0, double* tab)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL;
  const int nq = ipow(R.m, DIM - 1);
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq * NR) return;
  const int n = t % NR, q = t / NR, s = n / NL, i = n % NL;
  FaceGeo<DIM> G; G.init(m, cell0, face0);
  double jit = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d)
    if (d == G.axis) jit = (s == 0) ? 1.0 / (G.up_i[d] - G.lo_i[d]) : 1.0 / (G.up_o[d] - G.lo_o[d]);
  const double nrm0 = G.side ? 1.0 : -1.0;
  const double nrm = (s == 0) ? nrm0 : nrm0 * -1;
  const int sideval = (s == 0) ? G.side : 1 - G.side;      // geometryInInside: local[axis] = side; geometryInOutside: 1 - side
  double pos[3], w; quad_point<DIM - 1>(R, q, pos, w);
  double local[3]; int tt = 0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) local[d] = (d == G.axis) ? (double)sideval : pos[tt++];
  double js[3] = {0.0, 0.0, 0.0};
  qk_grad<DIM>(1, i, local, js);
  double dax = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; ++d) if (d == G.axis) dax = js[d];
  const double ga = 0.0 + jit * dax;
  const double gpn = 0.0 + ga * nrm;
  const double theta = 1;
  double* o = tab + (size_t)t * 4;
  o[0] = qk_phi<DIM>(1, i, local); o[1] = ga; o[2] = theta * 0.5 * gpn; o[3] = w;
}

// 1 = every face of the list has face number face0 and its inside cell lies in lattice plane `plane` along the face axis
template <int DIM>
__global__ void k_fd_planar(MeshDesc m, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces, int face0, int plane, int* bad)
{
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  int64_t c = fcell[f];
  int ijk[3]; ijk[0] = (int)(c % m.n[0]); c /= m.n[0]; ijk[1] = (int)(c % m.n[1]); ijk[2] = (int)(c / m.n[1]);
  if (fface[f] != face0 || ijk[face0 / 2] != plane) atomicExch(bad, 1);
}

// the parameter block: on-face values only, [side][q][on-face slot]
template <int DIM, int M> struct FdTab {
  static constexpr int NQ = (DIM == 3) ? M * M : M;
  static constexpr int NF = 1 << (DIM - 1);
  double phi[2][NQ][NF];
  double c[2][NQ][NF];
  double w[NQ];
};

template <int DIM, int AXIS> __host__ __device__ constexpr int fd_ion(int sideval, int a)
{
  return (a & ((1 << AXIS) - 1)) | (sideval << AXIS) | ((a >> AXIS) << (AXIS + 1));
}

template <int DIM, int M> struct FdSmem {
  static constexpr int NL = 1 << DIM, NR = 2 * NL, FPB = FD_NT / NR, NQ = FdTab<DIM, M>::NQ, PS = NR + 1;
  // per quadrature point (the q-stride is a compile-time constant: immediate offsets in the unrolled loop)
  struct PerQ {
    double2 prod[FPB * PS];        // {x_i phi_i, x_i ga_i}                face-major (stride PS = NR + 1: the faces of a warp hit different banks), side, node
    double2 mine[FD_NT];           // the same pair for the perturbed coefficient of thread t
    double2 base[FPB * 2];         // {u_s, gradu_s[axis]} base-line sums of side s
    double2 phic[NR];              // {phi_n, c_n} of node n (both sides), for the base-line row
    double2 pad[4];                // sizeof(PerQ) = 64 mod 128: the lanes of the base-sum phase (face k, point q) spread over bank quad k + 4 q
  };
  PerQ q[NQ];
  long long rowstart[FPB][NR];
  double ie[FPB], aoh[FPB];
  __device__ double* down(int fl) { return reinterpret_cast<double*>(q[0].mine) + fl * NR; }    // base-line rows; aliases q[0].mine after the main loop
};

template <int DIM, int AXIS, int SIDE, int M, int JS>
__device__ __forceinline__ void fd_body(const FdTab<DIM, M>& T, FdSmem<DIM, M>& S, const int fl, const int jj, const int t, const double ie, const double aoh,
                                        double (&r)[2 << DIM], double& rb)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL, NF = NL / 2, NQ = FdTab<DIM, M>::NQ;
  typedef typename FdSmem<DIM, M>::PerQ PerQ;
  constexpr int SV_SELF = (JS == 0) ? SIDE : 1 - SIDE;            // which lattice plane of the perturbed side's cell is the face
  // address of term i of the perturbed side's sums: this thread's own product for i == jj
  const char* addr[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) addr[i] = (i == jj) ? (const char*)&S.q[0].mine[t] : (const char*)&S.q[0].prod[fl * FdSmem<DIM, M>::PS + JS * NL + i];
  const char* abase_o = (const char*)&S.q[0].base[fl * 2 + (1 - JS)];
  const char* abase_s = (const char*)&S.q[0].base[fl * 2 + JS];
  const char* aphic = (const char*)&S.q[0].phic[JS * NL + jj];
  constexpr int nrm0_pos = SIDE;                                   // normal of side 0 = +e_axis iff SIDE == 1; side 1 has the opposite one
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    constexpr size_t QS = sizeof(PerQ);
    const double2 bo = *(const double2*)(abase_o + q * QS);
    double us = 0.0, gs = 0.0;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (((i >> AXIS) & 1) == SV_SELF) { const double2 a = *(const double2*)(addr[i] + q * QS); us += a.x; gs += a.y; }
      else gs += *(const double*)(addr[i] + q * QS + 8);
    }
    const double u0 = (JS == 0) ? us : bo.x, u1 = (JS == 0) ? bo.x : us;
    const double g0 = (JS == 0) ? gs : bo.y, g1 = (JS == 0) ? bo.y : gs;
    const double factor = T.w[q] * ie;
    const double jump[2] = {u0 - u1, u1 - u0};
    double mean = g0 + g1;
    mean *= 0.5;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double aj = aoh * jump[s];
      const int sv = (s == 0) ? SIDE : 1 - SIDE;
      const bool npos = (s == 0) ? (nrm0_pos != 0) : (nrm0_pos == 0);
#pragma unroll
      for (int a = 0; a < NF; ++a) {
        const double phi = T.phi[s][q][a];
        const double m1 = mean * phi;                 // |mgn * phi|: mgn = 0.0 + mean * (+-1)
        const double tt = T.c[s][q][a] * jump[s];
        double v = npos ? (-m1 - tt) : (m1 - tt);     // value = 0.0 - mgn*phi - c*jump
        v += aj * phi;
        const int ion = fd_ion<DIM, AXIS>(sv, a), ioff = ion ^ (1 << AXIS);
        r[s * NL + ion] += factor * v;
        r[s * NL + ioff] += factor * tt;              // opposite node: phi = +-0, c = -c_on  =>  value = c_on * jump
      }
    }
    // base line of row (JS, jj)
    {
      const double2 bs = *(const double2*)(abase_s + q * QS);
      const double2 pc = *(const double2*)(aphic + q * QS);
      const double jb = bs.x - bo.x;                                   // jump of this side: u_self - u_other
      double mb = (JS == 0) ? bs.y + bo.y : bo.y + bs.y;
      mb *= 0.5;
      const bool npos = (JS == 0) ? (nrm0_pos != 0) : (nrm0_pos == 0);
      const double m1 = mb * pc.x;
      const double tt = pc.y * jb;
      double v = npos ? (-m1 - tt) : (m1 - tt);
      v += (aoh * jb) * pc.x;
      rb += factor * v;
    }
  }
}

template <int DIM, int AXIS, int SIDE, int M>
__global__ void __launch_bounds__(FD_NT, 5) k_fd_coupling(const __grid_constant__ FdTab<DIM, M> T, MeshDesc m, const int32_t* __restrict__ fcell,
                                                       int64_t nfaces, double intensity, int analytic, double weight, double epsilon,
                                                       const double* __restrict__ gtab, const double* __restrict__ x, const int32_t* __restrict__ fdofs,
                                                       const int64_t* __restrict__ frowstart, const uint8_t* __restrict__ slotsT, double* val)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL, NQ = FdTab<DIM, M>::NQ;
  typedef FdSmem<DIM, M> Smem;
  constexpr int FPB = Smem::FPB, FW = 32 / NL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int js = w & 1, fl = (w >> 1) * FW + lane / NL, jj = lane % NL;
  const int n = js * NL + jj;
  const int64_t f = (int64_t)blockIdx.x * FPB + fl;
  const bool active = f < nfaces;
  double xj = 0.0;
  // the NR slot bytes of this thread's column (transposed table of the plan): one vector load, issued first
  typedef typename std::conditional<DIM == 3, uint4, uint2>::type SlotVec;
  SlotVec sl; memset(&sl, 0, sizeof(sl));
  if (active) {
    sl = __ldg(reinterpret_cast<const SlotVec*>(slotsT) + f * NR + n);
    const int32_t dof = __ldg(fdofs + f * NR + n);
    S.rowstart[fl][n] = __ldg(frowstart + f * NR + n);
    xj = analytic ? 0.0 : x[dof];
    if (n == 0) {
      FaceGeo<DIM> G; G.init(m, fcell[f], 2 * AXIS + SIDE);
      S.ie[fl] = G.integrationElement();
      S.aoh[fl] = intensity / G.cornerDistance01();
    }
  }
  double delta = 1.0, xp = 1.0;
  if (!analytic) { delta = epsilon * (1.0 + fabs(xj)); xp = xj + delta; }
  // phase 1: the products of this node at every quadrature point
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double* g = gtab + (size_t)(q * NR + n) * 4;
    const double ph = __ldg(g), ga = __ldg(g + 1);
    S.q[q].prod[fl * Smem::PS + n] = make_double2(xj * ph, xj * ga);
    S.q[q].mine[t] = make_double2(xp * ph, xp * ga);
    if (fl == 0) S.q[q].phic[n] = make_double2(ph, __ldg(g + 2));
  }
  __syncthreads();
  // base-line sums of side js of this face at the quadrature points q = jj, jj + NL, ... (sequential, reference order)
  {
    const int sv_self = js ? 1 - SIDE : SIDE;                      // which lattice plane of that side's cell is the face
    for (int q = jj; q < NQ; q += NL) {
      const double2* p = S.q[q].prod + fl * Smem::PS + js * NL;
      double us = 0.0, gs = 0.0;
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const double2 a = p[i];
        if (((i >> AXIS) & 1) == sv_self) us += a.x;
        gs += a.y;
      }
      S.q[q].base[fl * 2 + js] = make_double2(us, gs);
    }
  }
  __syncthreads();
  double r[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) r[i] = 0.0;
  double rb = 0.0;
  const double ie = S.ie[active ? fl : 0], aoh = S.aoh[active ? fl : 0];
  if (js == 0) fd_body<DIM, AXIS, SIDE, M, 0>(T, S, fl, jj, t, ie, aoh, r, rb);
  else fd_body<DIM, AXIS, SIDE, M, 1>(T, S, fl, jj, t, ie, aoh, r, rb);
  __syncthreads();                                                   // every thread is past the main loop: q[0].mine is free
  S.down(fl)[n] = rb;
  __syncthreads();
  if (!active) return;
  const double inv = 1.0 / delta;
  const uint8_t* slb = reinterpret_cast<const uint8_t*>(&sl);         // register-resident after unrolling
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    double c = r[i];
    if (!analytic) {
      // (up - down) / delta by Markstein's sequence on the reciprocal formed once per thread: q0 = RN(a y), q = RN(q0 + (a - delta q0) y) —
      // the correctly rounded quotient (y = RN(1 / delta), q0 within an ulp); an exact zero stays an exact zero
      const double a = r[i] - S.down(fl)[i];
      const double q0 = a * inv;
      c = fma(fma(-delta, q0, a), inv, q0);
    }
    // an entry this column does not influence (both nodes off the face: the row sees neither the jump nor the mean gradient change) is an
    // exact +0.0: nothing to add
    if (c != 0.0) atomicAdd(&val[S.rowstart[fl][i] + slb[i]], weight * c);         // mat_ss / mat_ns (j self), mat_sn / mat_nn (j neighbour)
  }
}

// plan tables: per (face, node) the DOF and the start of its row; the slot table transposed to [face][column][row]
template <int DIM>
__global__ void k_fd_plan_tables(MeshDesc m, SideMap Ls, SideMap Rs, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces,
                                 const int64_t* __restrict__ rowptr, const uint8_t* __restrict__ slots, int32_t* fdofs, int64_t* frowstart, uint8_t* slotsT)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nfaces * NR) return;
  const int64_t f = t / NR; const int n = (int)(t % NR);
  FaceGeo<DIM> G; G.init(m, fcell[f], fface[f]);
  int32_t ds[NL], dn[NL];
  cell_dofs_q1<DIM>(Ls, G.ijk, ds); cell_dofs_q1<DIM>(Rs, G.nijk, dn);
  int32_t dof = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) { if (n == i) dof = ds[i]; if (n == NL + i) dof = dn[i]; }
  fdofs[t] = dof; frowstart[t] = rowptr[dof];
  for (int i = 0; i < NR; ++i) slotsT[t * NR + i] = slots[f * (NR * NR) + i * NR + n];
}

struct FdPlan {
  int usable = 0;                 // 0: the face list is not planar / the tables violate an identity -> staged kernel
  int dim = 0, axis = 0, side = 0, m = 0;
  double* d_tab = nullptr;
  std::vector<double> h_tab;      // nq * NR * 4 doubles
  int32_t* d_dofs = nullptr; int64_t* d_rowstart = nullptr; uint8_t* d_slotsT = nullptr;
};

void fd_plan_free(FdPlan* p) { if (p) { dfree(p->d_tab); dfree(p->d_dofs); dfree(p->d_rowstart); dfree(p->d_slotsT); delete p; } }

template <int DIM> static FdPlan* fd_plan_build(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, const FaceList& fl, const Matrix& Mx, const uint8_t* slots)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL;
  FdPlan* plan = new FdPlan();
  plan->dim = DIM; plan->m = P.rule.m;
  const int nq = (DIM == 3) ? P.rule.m * P.rule.m : P.rule.m;
  if (P.rule.m < 2 || P.rule.m > 4 || fl.n == 0) return plan;
  int32_t cell0 = 0; int8_t face0 = 0;
  MDB_CUDA(cudaMemcpyAsync(&cell0, fl.d_cell, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&face0, fl.d_face, sizeof(int8_t), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  const MeshDesc& m = c->mesh;
  plan->axis = face0 / 2; plan->side = face0 % 2;
  int64_t cc = cell0; int ijk[3]; ijk[0] = (int)(cc % m.n[0]); cc /= m.n[0]; ijk[1] = (int)(cc % m.n[1]); ijk[2] = (int)(cc / m.n[1]);
  int* d_bad = dalloc<int>(1);
  MDB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
  MDB_LAUNCH(c, k_fd_planar<DIM>, cdiv(fl.n, 256), 256, 0, m, fl.d_cell, fl.d_face, fl.n, (int)face0, ijk[plan->axis], d_bad);
  plan->d_tab = dalloc<double>((size_t)nq * NR * 4);
  MDB_LAUNCH(c, k_fd_tables<DIM>, cdiv(nq * NR, 128), 128, 0, m, P.rule, cell0, (int)face0, plan->d_tab);
  plan->h_tab.resize((size_t)nq * NR * 4);
  int bad = 0;
  MDB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(plan->h_tab.data(), plan->d_tab, plan->h_tab.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  dfree(d_bad);
  if (bad) return plan;
  // the two identities the kernel relies on, checked on the bits
  for (int q = 0; q < nq; ++q)
    for (int s = 0; s < 2; ++s) {
      const int sv = (s == 0) ? plan->side : 1 - plan->side;
      for (int i = 0; i < NL; ++i) {
        if (((i >> plan->axis) & 1) != sv) continue;
        const double* on = &plan->h_tab[(size_t)(q * NR + s * NL + i) * 4];
        const double* off = &plan->h_tab[(size_t)(q * NR + s * NL + (i ^ (1 << plan->axis))) * 4];
        if (!(off[0] == 0.0) || !(off[2] == -on[2]) || !(off[1] == -on[1])) return plan;
      }
    }
  plan->d_dofs = dalloc<int32_t>((size_t)fl.n * NR); plan->d_rowstart = dalloc<int64_t>((size_t)fl.n * NR); plan->d_slotsT = dalloc<uint8_t>((size_t)fl.n * NR * NR);
  MDB_LAUNCH(c, k_fd_plan_tables<DIM>, cdiv(fl.n * NR, 256), 256, 0, m, Ls, Rs, fl.d_cell, fl.d_face, fl.n, Mx.d_rowptr, slots, plan->d_dofs, plan->d_rowstart,
             plan->d_slotsT);
  plan->usable = 1;
  return plan;
}

template <int DIM, int AXIS, int SIDE, int M>
static void fd_launch(Ctx* c, const FdPlan& plan, const CouplingParams& P, const FaceList& fl, Matrix& Mx, double weight, double eps, const double* d_x)
{
  constexpr int NL = 1 << DIM, NR = 2 * NL, NF = NL / 2;
  typedef FdTab<DIM, M> Tab;
  typedef FdSmem<DIM, M> Smem;
  Tab T;
  for (int q = 0; q < Tab::NQ; ++q) {
    for (int s = 0; s < 2; ++s) {
      const int sv = (s == 0) ? SIDE : 1 - SIDE;
      for (int a = 0; a < NF; ++a) {
        const double* e = &plan.h_tab[(size_t)(q * NR + s * NL + fd_ion<DIM, AXIS>(sv, a)) * 4];
        T.phi[s][q][a] = e[0]; T.c[s][q][a] = e[2];
      }
    }
    T.w[q] = plan.h_tab[(size_t)(q * NR) * 4 + 3];
  }
  // MDB_FD_BLOCKS_PER_SM = k bounds the kernel's footprint to k resident blocks per SM by padding its dynamic shared memory: beside
  // the per-thread-store fill (MDB_FILL_NO_TMA) the FP64-bound faces then leave the SM's warp slots to the stores
  static const int per_sm = getenv("MDB_FD_BLOCKS_PER_SM") ? atoi(getenv("MDB_FD_BLOCKS_PER_SM")) : 0;
  size_t smem_bytes = sizeof(Smem);
  if (per_sm >= 1 && per_sm <= 8) smem_bytes = std::max(smem_bytes, (size_t)(224 * 1024 / per_sm - 1024));
  static bool configured = false;
  if (!configured) {
    MDB_CUDA(cudaFuncSetAttribute(k_fd_coupling<DIM, AXIS, SIDE, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    configured = true;
  }
  MDB_LAUNCH(c, (k_fd_coupling<DIM, AXIS, SIDE, M>), cdiv(fl.n, Smem::FPB), FD_NT, smem_bytes, T, c->mesh, fl.d_cell, fl.n, P.intensity,
             (int)(P.jac_mode == MDB_JAC_ANALYTIC), weight, eps, plan.d_tab, d_x, plan.d_dofs, plan.d_rowstart, plan.d_slotsT, Mx.d_val);
}

template <int DIM, int AXIS, int SIDE, typename... A> static void fd_dispatch_m(int m, A&&... a)
{
  if (m == 2) fd_launch<DIM, AXIS, SIDE, 2>(a...);
  else if (m == 3) fd_launch<DIM, AXIS, SIDE, 3>(a...);
  else fd_launch<DIM, AXIS, SIDE, 4>(a...);
}
template <int DIM, int AXIS, typename... A> static void fd_dispatch_side(int side, int m, A&&... a)
{
  if (side) fd_dispatch_m<DIM, AXIS, 1>(m, a...); else fd_dispatch_m<DIM, AXIS, 0>(m, a...);
}

// returns false when the fast kernel does not apply (the caller then runs the staged kernel)
bool jacobian_coupling_fd_fast(Ctx* c, const CouplingParams& P, const SideMap& Ls, const SideMap& Rs, const FaceList& fl, Matrix& M, std::pair<int, int> key,
                               const uint8_t* slots, double weight, double eps, const double* d_x)
{
  static const bool off = getenv("MDB_FD_STAGED") != nullptr;          // A/B switch: always the staged kernel
  if (off || P.op != MDB_OP_CVCF_COUPLING) return false;
  const int dim = c->mesh.dim;
  auto it = M.fd_plans.find(key);
  if (it == M.fd_plans.end()) {
    FdPlan* p = (dim == 2) ? fd_plan_build<2>(c, P, Ls, Rs, fl, M, slots) : fd_plan_build<3>(c, P, Ls, Rs, fl, M, slots);
    it = M.fd_plans.insert(std::make_pair(key, p)).first;
  }
  const FdPlan& plan = *it->second;
  if (!plan.usable || plan.m != P.rule.m) return false;
  if (dim == 2) {
    if (plan.axis == 0) fd_dispatch_side<2, 0>(plan.side, plan.m, c, plan, P, fl, M, weight, eps, d_x);
    else fd_dispatch_side<2, 1>(plan.side, plan.m, c, plan, P, fl, M, weight, eps, d_x);
  } else {
    if (plan.axis == 0) fd_dispatch_side<3, 0>(plan.side, plan.m, c, plan, P, fl, M, weight, eps, d_x);
    else if (plan.axis == 1) fd_dispatch_side<3, 1>(plan.side, plan.m, c, plan, P, fl, M, weight, eps, d_x);
    else fd_dispatch_side<3, 2>(plan.side, plan.m, c, plan, P, fl, M, weight, eps, d_x);
  }
  return true;
}

} // namespace mdb
