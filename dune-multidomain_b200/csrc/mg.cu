// mg.cu — MDB_PRECOND_MG: a geometric multigrid V-cycle as preconditioner of the Krylov solvers of solver.cu, for the structured
// Q1 multidomain systems (the north star's "CG/BiCGSTAB solver ... using SpMV and fused dot/axpy kernels").
//
// The reference's solver backends precondition with sequential ISTL sweeps (SeqSSOR / SeqILU0, precond.cu reproduces them operation
// for operation); their iteration counts grow like 1/h and a sweep is a dependency chain, so on a GPU Jacobi-CG wins in time although
// it needs ~1000 iterations at 256^3.  The bandwidth-bound alternative that removes the 1/h is multigrid, and on this path every
// ingredient already exists as a streamed kernel:
//   coarse operators   re-discretisation: the SAME assembly path on the mesh with half the cells per axis (a coarse context is built by
//                      replaying the recorded setup: spaces, subproblems, couplings, constraints, grid operators, and the log of
//                      (grid operator, weight) Jacobian calls that produced the fine matrix — so one-step matrices M + dt K coarsen too)
//   transfers          per child space, trilinear interpolation between the child's lattices of two levels (the two subdomains keep
//                      their own coefficients on the interface, exactly as in the DOF map); restriction = its transpose; constrained
//                      coefficients are neither corrected nor restricted to
//   smoother           Chebyshev polynomial in D^-1 A on [lambda_max / 4, 1.1 lambda_max] (lambda_max by a power iteration at set-up):
//                      symmetric, no dependency chain, one SpMV per degree — the V-cycle is an SPD operator, valid inside CG
//   coarsest level     the same smoother with a wider interval and more steps (a few dozen coefficients)
// Q1 children (also on a slab partition) or QkDG children (nested spaces: transfers by evaluation of the coarse cell's basis), meshes
// whose cell counts and subdomain marking coarsen (even counts, uniform 2^d blocks).
#include "common.cuh"
#include <cmath>

struct mdb_ctx : mdb::Ctx {};

namespace mdb {

namespace {

struct MgLevel {
  mdb_ctx* ctx = nullptr; bool owned = false; int mat = -1;
  int64_t n = 0;
  double *x = nullptr, *b = nullptr, *r = nullptr, *d = nullptr, *t = nullptr, *dinv = nullptr;
  double lmax = 0.0;
};
struct MgHier {
  std::vector<MgLevel> L;
  std::vector<std::pair<int, double>> log;      // the assembly log the coarse matrices were built from
  int degree = 2, coarse_degree = 16;
  double* d_red = nullptr;
};

// `shift`: local fine index of the first fine layer under local coarse layer 0 along the partition axis (0 without a partition; on a
// rank with a ghost layer below it is -1: only the upper of the two fine layers under the coarse ghost layer is local)
__global__ void k_mg_coarsen_sd(MeshDesc mf, MeshDesc mc, int shift, const uint8_t* __restrict__ sdf, uint8_t* sdc, int* bad)
{
  const int64_t cidx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (cidx >= mc.ncells()) return;
  int I[3]; int64_t r = cidx;
  I[0] = (int)(r % mc.n[0]); r /= mc.n[0]; I[1] = (int)(r % mc.n[1]); I[2] = (int)(r / mc.n[1]);
  const int pa = mf.dim - 1;
  uint8_t v = 0; bool first = true;
  for (int dz = 0; dz < (mf.dim > 2 ? 2 : 1); ++dz)
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        int F[3] = {2 * I[0] + dx, 2 * I[1] + dy, mf.dim > 2 ? 2 * I[2] + dz : 0};
        F[pa] += shift;
        if (F[pa] < 0 || F[pa] >= mf.n[pa]) continue;
        const int64_t f = F[0] + (int64_t)mf.n[0] * (F[1] + (int64_t)mf.n[1] * F[2]);
        const uint8_t s = sdf[f];
        if (first) { v = s; first = false; } else if (s != v) *bad = 1;
      }
  sdc[cidx] = v;
}
__global__ void k_mg_dinv(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ diag, const double* __restrict__ val, double* dinv)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 1.0;
  if (diag[i] >= 0) { const double a = val[rowptr[i] + diag[i]]; d = a != 0.0 ? 1.0 / a : 1.0; }
  dinv[i] = d;
}
struct MgChild { int64_t off_f, off_c, size_f, size_c; const int32_t *d2l_f, *l2d_f, *d2l_c, *l2d_c; int nf[3], nc[3]; int gf[3], gc[3]; int k, dg; };   // gf / gc: global lattice index of local plane 0 (partition)

// x_f += P x_c on the coefficients of one child (not on constrained ones)
__global__ void k_mg_prolong(MgChild C, int dim, const uint8_t* __restrict__ con_f, const double* __restrict__ xc, double* xf)
{
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= C.size_f) return;
  if (con_f[C.off_f + g]) return;
  int64_t p = C.d2l_f[g];
  int i[3] = {0, 0, 0};
  i[0] = (int)(p % C.nf[0]); p /= C.nf[0]; i[1] = (int)(p % C.nf[1]); i[2] = (int)(p / C.nf[1]);
  int lo[3], cnt[3]; double w[3];
  for (int d = 0; d < 3; ++d) {
    if (d >= dim) { lo[d] = 0; cnt[d] = 1; w[d] = 1.0; continue; }
    const int ig = i[d] + C.gf[d];                                   // global lattice index
    if ((ig & 1) == 0) { lo[d] = (ig >> 1) - C.gc[d]; cnt[d] = 1; w[d] = 1.0; } else { lo[d] = ((ig - 1) >> 1) - C.gc[d]; cnt[d] = 2; w[d] = 0.5; }
  }
  double acc = 0.0;
  for (int c = 0; c < cnt[2]; ++c)
    for (int b = 0; b < cnt[1]; ++b)
      for (int a = 0; a < cnt[0]; ++a) {
        if (lo[0] + a < 0 || lo[0] + a >= C.nc[0] || lo[1] + b < 0 || lo[1] + b >= C.nc[1] || lo[2] + c < 0 || lo[2] + c >= C.nc[2]) continue;   // a ghost coefficient whose parent is not local
        const int64_t pc = (lo[0] + a) + (int64_t)C.nc[0] * ((lo[1] + b) + (int64_t)C.nc[1] * (lo[2] + c));
        const int32_t dc = C.l2d_c[pc];
        if (dc >= 0) acc += w[0] * w[1] * w[2] * xc[C.off_c + dc];
      }
  xf[C.off_f + g] += acc;
}
// r_c = P^T r_f on the coefficients of one child; constrained coarse coefficients get 0
__global__ void k_mg_restrict(MgChild C, int dim, const uint8_t* __restrict__ con_f, const uint8_t* __restrict__ con_c, const double* __restrict__ rf, double* rc)
{
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= C.size_c) return;
  if (con_c[C.off_c + g]) { rc[C.off_c + g] = 0.0; return; }
  int64_t p = C.d2l_c[g];
  int I[3] = {0, 0, 0};
  I[0] = (int)(p % C.nc[0]); p /= C.nc[0]; I[1] = (int)(p % C.nc[1]); I[2] = (int)(p / C.nc[1]);
  double acc = 0.0;
  for (int c = (dim > 2 ? -1 : 0); c <= (dim > 2 ? 1 : 0); ++c)
    for (int b = -1; b <= 1; ++b)
      for (int a = -1; a <= 1; ++a) {
        const int fi = 2 * (I[0] + C.gc[0]) + a - C.gf[0], fj = 2 * (I[1] + C.gc[1]) + b - C.gf[1], fk = 2 * (I[2] + C.gc[2]) + c - C.gf[2];
        if (fi < 0 || fi >= C.nf[0] || fj < 0 || fj >= C.nf[1] || fk < 0 || fk >= C.nf[2]) continue;
        const int32_t df = C.l2d_f[fi + (int64_t)C.nf[0] * (fj + (int64_t)C.nf[1] * fk)];
        if (df < 0 || con_f[C.off_f + df]) continue;
        const double w = (a ? 0.5 : 1.0) * (b ? 0.5 : 1.0) * (c ? 0.5 : 1.0);
        acc += w * rf[C.off_f + df];
      }
  rc[C.off_c + g] = acc;
}

// QkDG children: the spaces of two levels are nested (a coarse cell's polynomial restricted to its 2^d children), so the
// prolongation is the evaluation of the coarse cell's Lagrange basis at the fine cell's nodes and the restriction its transpose.
// Lattice of a QkDG child: cell c, node a  <->  lattice point (k + 1) c + a per axis; nodes equidistant.
__device__ inline double mg_lagrange(int k, int a, double xi)
{
  double v = 1.0;
  for (int m = 0; m <= k; ++m) if (m != a) v *= (xi * k - m) / (double)(a - m);
  return v;
}
__global__ void k_mg_prolong_dg(MgChild C, int dim, const double* __restrict__ xc, double* xf)
{
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= C.size_f) return;
  const int k = C.k, k1 = C.k + 1;
  int64_t p = C.d2l_f[g];
  int pf[3] = {0, 0, 0};
  pf[0] = (int)(p % C.nf[0]); p /= C.nf[0]; pf[1] = (int)(p % C.nf[1]); pf[2] = (int)(p / C.nf[1]);
  int I[3] = {0, 0, 0}; double w[3][3] = {{1, 0, 0}, {1, 0, 0}, {1, 0, 0}};
  for (int d = 0; d < dim; ++d) {
    const int i = pf[d] / k1, a = pf[d] % k1;
    I[d] = i >> 1;
    const double xi = 0.5 * ((i & 1) + (double)a / k);
    for (int b = 0; b <= k; ++b) w[d][b] = mg_lagrange(k, b, xi);
  }
  double acc = 0.0;
  for (int c = 0; c <= (dim > 2 ? k : 0); ++c)
    for (int b = 0; b <= k; ++b)
      for (int a = 0; a <= k; ++a) {
        const int64_t pc = (k1 * I[0] + a) + (int64_t)C.nc[0] * ((k1 * I[1] + b) + (int64_t)C.nc[1] * (dim > 2 ? k1 * I[2] + c : 0));
        const int32_t dc = C.l2d_c[pc];
        if (dc >= 0) acc += w[0][a] * w[1][b] * (dim > 2 ? w[2][c] : 1.0) * xc[C.off_c + dc];
      }
  xf[C.off_f + g] += acc;
}
__global__ void k_mg_restrict_dg(MgChild C, int dim, const double* __restrict__ rf, double* rc)
{
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= C.size_c) return;
  const int k = C.k, k1 = C.k + 1;
  int64_t p = C.d2l_c[g];
  int pc[3] = {0, 0, 0};
  pc[0] = (int)(p % C.nc[0]); p /= C.nc[0]; pc[1] = (int)(p % C.nc[1]); pc[2] = (int)(p / C.nc[1]);
  int I[3], A[3];
  for (int d = 0; d < 3; ++d) { I[d] = pc[d] / k1; A[d] = pc[d] % k1; }
  double acc = 0.0;
  const int nz = dim > 2 ? 2 * k1 : 1;
  for (int fz = 0; fz < nz; ++fz) {
    const double wz = dim > 2 ? mg_lagrange(k, A[2], 0.5 * ((fz / k1) + (double)(fz % k1) / k)) : 1.0;
    for (int fy = 0; fy < 2 * k1; ++fy) {
      const double wy = mg_lagrange(k, A[1], 0.5 * ((fy / k1) + (double)(fy % k1) / k));
      for (int fx = 0; fx < 2 * k1; ++fx) {
        const double wx = mg_lagrange(k, A[0], 0.5 * ((fx / k1) + (double)(fx % k1) / k));
        // fine lattice point: cell 2 I + (f / k1), node f % k1  ->  k1 (2 I) + f
        const int64_t lf = (2 * k1 * I[0] + fx) + (int64_t)C.nf[0] * ((2 * k1 * I[1] + fy) + (int64_t)C.nf[1] * (dim > 2 ? 2 * k1 * I[2] + fz : 0));
        const int32_t df = C.l2d_f[lf];
        if (df >= 0) acc += wx * wy * wz * rf[C.off_f + df];
      }
    }
  }
  rc[C.off_c + g] = acc;
}
// Chebyshev smoothing steps (see mg_smooth)
__global__ void k_mg_cheb_start(int64_t n, const double* __restrict__ b, const double* __restrict__ ax, const double* __restrict__ dinv, double inv_theta,
                                double* r, double* d, double* x, int zero_guess)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = zero_guess ? b[i] : b[i] - ax[i];
    const double di = inv_theta * dinv[i] * ri;
    r[i] = ri; d[i] = di;
    x[i] = zero_guess ? di : x[i] + di;
  }
}
__global__ void k_mg_cheb_step(int64_t n, const double* __restrict__ ad, const double* __restrict__ dinv, double c1, double c2, double* r, double* d, double* x)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i] - ad[i];
    const double di = c1 * d[i] + c2 * dinv[i] * ri;
    r[i] = ri; d[i] = di; x[i] += di;
  }
}
__global__ void k_mg_sub(int64_t n, const double* __restrict__ b, const double* __restrict__ ax, double* r)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) r[i] = b[i] - ax[i];
}
__global__ void k_mg_scale_dinv(int64_t n, const double* __restrict__ dinv, const double* __restrict__ av, double* w)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] = dinv[i] * av[i];
}
__global__ void k_mg_init_vec(int64_t n, double* v)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = 1.0 + 0.5 * sin(0.7 * (double)i);
}
__global__ void k_mg_scale(int64_t n, double a, double* v)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] *= a;
}
__global__ void k_mg_norm2(int64_t n, const double* __restrict__ v, const uint8_t* __restrict__ owned, double* out)
{
  __shared__ double sh[8];
  double a = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) if (!owned || owned[i]) a += v[i] * v[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i]; atomicAdd(out, s); }
}

// y = A x on a level; on a partition the ghost entries of x are refreshed first (x must be writable)
void level_spmv(Ctx* lc, Matrix& M, double* x, double* y)
{
  if (lc->mesh.partitioned && lc->nranks > 1) spmv_partitioned(lc, M, x, y); else spmv(lc, M, x, y);
}
int vgrid(int64_t n) { const int g = cdiv(n, 256); return g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8; }

// 2-norm over the owned rows of all ranks (lc: the level's context)
double norm2(Ctx* c, Ctx* lc, MgHier* H, int64_t n, const double* v)
{
  MDB_CUDA(cudaMemsetAsync(H->d_red, 0, sizeof(double), c->stream));
  MDB_LAUNCH(c, k_mg_norm2, vgrid(n), 256, 0, n, v, lc->d_owned, H->d_red);
  if (lc->nranks > 1) allreduce_scalars(lc, H->d_red, 1);
  double h = 0.0;
  MDB_CUDA(cudaMemcpyAsync(&h, H->d_red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  return std::sqrt(h);
}

#define MG_CALL(ctx, expr, what)                                                                                        \
  do { const int rc__ = (expr); if (rc__ < 0) throw Error(rc__, std::string("multigrid set-up: ") + what + ": " + mdb_last_error(ctx)); } while (0)

// build the context of the next coarser level by replaying the set-up of `f`; returns nullptr when the level does not coarsen
// all ranks agree: the sum of the flags over the ranks equals the number of ranks
bool all_ranks(Ctx* f, bool mine)
{
  if (f->nranks <= 1) return mine;
  double h = mine ? 1.0 : 0.0;
  double* d = dalloc<double>(1);
  MDB_CUDA(cudaMemcpyAsync(d, &h, sizeof(double), cudaMemcpyHostToDevice, f->stream));
  allreduce_scalars(f, d, 1);
  MDB_CUDA(cudaMemcpyAsync(&h, d, sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  MDB_CUDA(cudaStreamSynchronize(f->stream));
  cudaFree(d);
  return h > f->nranks - 0.5;
}

mdb_ctx* coarsen(Ctx* f)
{
  const MeshDesc& m = f->mesh;
  const int pa = m.dim - 1;
  int64_t nc[3] = {1, 1, 1}, gnc[3] = {1, 1, 1}, offc[3] = {0, 0, 0}; double lower[3], upper[3], glo[3], gup[3];
  bool ok = true; int shift = 0;
  for (int d = 0; d < m.dim; ++d) {
    if (m.gn[d] % 2 != 0 || m.gn[d] < 4) ok = false;
    gnc[d] = m.gn[d] / 2; glo[d] = m.glower[d]; gup[d] = m.glower[d] + m.h[d] * m.gn[d];
    if (m.partitioned && d == pa) {
      // owned cell layers [c0, c1) of this slab; the coarse slab owns [c0/2, c1/2) and keeps one coarse ghost layer below
      const bool first = m.off[d] == 0;
      const int c0 = m.off[d] + (first ? 0 : 1), c1 = m.off[d] + m.n[d];
      // at least FOUR owned layers: the coarse slab then owns two.  A coarse first slab of one layer would put the second rank's ghost
      // layer at global layer 0, and mdb_mesh_partition tells the first rank by cell offset 0 (found on 4 B200s with 16 layers:
      // 4 -> 2 -> 1 layers per rank, "lower neighbour does not match the partition" at the third level)
      if (c0 % 2 != 0 || c1 % 2 != 0 || c1 - c0 < 4) ok = false;
      offc[d] = c0 / 2 - (first ? 0 : 1); nc[d] = c1 / 2 - offc[d];
      shift = (int)(2 * offc[d] - m.off[d]);
    } else {
      if (m.n[d] % 2 != 0) ok = false;
      offc[d] = m.off[d] / 2; nc[d] = m.n[d] / 2;
      if (m.off[d] % 2 != 0) ok = false;
    }
    lower[d] = glo[d] + (double)offc[d] * 2.0 * m.h[d]; upper[d] = glo[d] + (double)(offc[d] + nc[d]) * 2.0 * m.h[d];
  }
  if (!all_ranks(f, ok)) return nullptr;
  mdb_ctx* c = mdb_create(f->device);
  MDB_REQUIRE(c, MDB_ECUDA, "multigrid set-up: cannot create a coarse context");
  try {
    MG_CALL(c, mdb_mesh_structured(c, m.dim, nc, lower, upper), "mesh");
    if (m.partitioned) MG_CALL(c, mdb_mesh_partition(c, gnc, offc, glo, gup), "partition");
    int* d_bad = dalloc<int>(1);
    MDB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), f->stream));
    MDB_LAUNCH(f, k_mg_coarsen_sd, cdiv(c->mesh.ncells(), 256), 256, 0, m, c->mesh, shift, f->d_cell_sd, c->d_cell_sd, d_bad);
    int bad = 0;
    MDB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
    MDB_CUDA(cudaStreamSynchronize(f->stream));
    cudaFree(d_bad);
    if (!all_ranks(f, bad == 0)) { mdb_destroy(c); return nullptr; }       // the subdomain marking is not a union of 2^d blocks: this level is the coarsest
    // the coarse level reduces over the same ranks: it borrows the communicator
    c->nccl_comm = f->nccl_comm; c->comm_borrowed = true; c->nranks = f->nranks; c->rank = f->rank;
    c->sring = f->sring; c->sring_borrowed = f->sring != nullptr;
    c->have_sd = true;
    for (const Child& ch : f->children) MG_CALL(c, mdb_add_space(c, ch.fe_kind, ch.subdomain), "space");
    for (const SubProblemDesc& sp : f->sps) {
      const void* p = sp.op_id == MDB_OP_POISSON ? (const void*)&sp.poisson : sp.op_id == MDB_OP_CONVDIFF_DG ? (const void*)&sp.cddg : (const void*)&sp.l2;
      const size_t nb = sp.op_id == MDB_OP_POISSON ? sizeof(sp.poisson) : sp.op_id == MDB_OP_CONVDIFF_DG ? sizeof(sp.cddg) : sizeof(sp.l2);
      MG_CALL(c, mdb_add_subproblem(c, sp.op_id, p, nb, sp.cond_kind, sp.mask, sp.children.data(), (int)sp.children.size()), "subproblem");
    }
    for (const CouplingDesc& cp : f->cps) {
      const void* p = cp.op_id == MDB_OP_CVCF_COUPLING ? (const void*)&cp.cvcf : (const void*)&cp.propflow;
      const size_t nb = cp.op_id == MDB_OP_CVCF_COUPLING ? sizeof(cp.cvcf) : sizeof(cp.propflow);
      MG_CALL(c, mdb_add_coupling(c, cp.op_id, p, nb, cp.local_sp, cp.remote_sp, MDB_JAC_ANALYTIC), "coupling");     // linear couplings: the exact derivative
    }
    int64_t nd = 0;
    MG_CALL(c, mdb_finalize(c, &nd), "finalize");
    if (!f->con_sps.empty()) { int64_t ncon = 0; MG_CALL(c, mdb_constraints_assemble(c, f->con_sps.data(), f->con_bcs.data(), (int)f->con_sps.size(), &ncon), "constraints"); }
    mdb_set_time(c, f->time);
    for (const GridOp& g : f->gos) MG_CALL(c, mdb_gridoperator_create(c, g.sps.data(), (int)g.sps.size(), g.cps.data(), (int)g.cps.size()), "grid operator");
  } catch (...) { mdb_destroy(c); throw; }
  return c;
}

void level_alloc(MgLevel& l, bool top)
{
  l.r = dalloc<double>(l.n); l.d = dalloc<double>(l.n); l.t = dalloc<double>(l.n); l.dinv = dalloc<double>(l.n);
  if (!top) { l.x = dalloc<double>(l.n); l.b = dalloc<double>(l.n); }
}
void level_free(MgLevel& l, bool top)
{
  if (l.r) cudaFree(l.r); if (l.d) cudaFree(l.d); if (l.t) cudaFree(l.t); if (l.dinv) cudaFree(l.dinv);
  if (!top) { if (l.x) cudaFree(l.x); if (l.b) cudaFree(l.b); }
  if (l.owned && l.ctx) mdb_destroy(l.ctx);
}

// assemble the level's matrix from the log, extract D^-1, estimate lambda_max(D^-1 A)
void level_numeric(Ctx* c, MgHier* H, size_t li)
{
  MgLevel& l = H->L[li];
  Ctx* lc = l.ctx;
  StreamScope scope(lc, c->stream);
  Matrix& M = lc->mats[l.mat];
  if (li > 0) {
    MDB_CUDA(cudaMemsetAsync(l.x, 0, l.n * sizeof(double), c->stream));
    bool first = true;
    for (const auto& e : H->log) { MG_CALL(l.ctx, mdb_jacobian(l.ctx, e.first, l.mat, e.second, l.x, first ? 1 : 0), "coarse Jacobian"); first = false; }
  }
  MDB_LAUNCH(c, k_mg_dinv, cdiv(l.n, 256), 256, 0, l.n, M.d_rowptr, M.d_diag, M.d_val, l.dinv);
  // power iteration for lambda_max(D^-1 A)
  MDB_LAUNCH(c, k_mg_init_vec, vgrid(l.n), 256, 0, l.n, l.d);
  double nv = norm2(c, lc, H, l.n, l.d), lam = 1.0;
  for (int it = 0; it < 12; ++it) {
    MDB_LAUNCH(c, k_mg_scale, vgrid(l.n), 256, 0, l.n, 1.0 / nv, l.d);
    level_spmv(lc, M, l.d, l.t);
    MDB_LAUNCH(c, k_mg_scale_dinv, vgrid(l.n), 256, 0, l.n, l.dinv, l.t, l.d);
    nv = norm2(c, lc, H, l.n, l.d);
    lam = nv;
  }
  l.lmax = lam;
}

// `degree` Chebyshev steps for A x = b in the D^-1-scaled operator on [lo, hi]
void mg_smooth(Ctx* c, MgLevel& l, const double* b, double* x, bool zero_guess, int degree, double lo, double hi)
{
  Ctx* lc = l.ctx;
  Matrix& M = lc->mats[l.mat];
  const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
  double rho_old = 1.0 / sigma;
  if (!zero_guess) level_spmv(lc, M, x, l.t);
  MDB_LAUNCH(c, k_mg_cheb_start, vgrid(l.n), 256, 0, l.n, b, l.t, l.dinv, 1.0 / theta, l.r, l.d, x, zero_guess ? 1 : 0);
  for (int k = 1; k < degree; ++k) {
    const double rho = 1.0 / (2.0 * sigma - rho_old);
    level_spmv(lc, M, l.d, l.t);
    MDB_LAUNCH(c, k_mg_cheb_step, vgrid(l.n), 256, 0, l.n, l.t, l.dinv, rho * rho_old, 2.0 * rho / delta, l.r, l.d, x);
    rho_old = rho;
  }
}

MgChild child_pair(const Ctx* f, const Ctx* cc, size_t k)
{
  const Child& a = f->children[k]; const Child& b = cc->children[k];
  MgChild C; C.off_f = a.offset; C.off_c = b.offset; C.size_f = a.size; C.size_c = b.size;
  C.d2l_f = a.d_dof2lat; C.l2d_f = a.d_lat2dof; C.d2l_c = b.d_dof2lat; C.l2d_c = b.d_lat2dof;
  for (int d = 0; d < 3; ++d) { C.nf[d] = a.lat_n[d]; C.nc[d] = b.lat_n[d]; C.gf[d] = d < f->mesh.dim ? f->mesh.off[d] : 0; C.gc[d] = d < cc->mesh.dim ? cc->mesh.off[d] : 0; }
  C.k = a.k; C.dg = a.dg ? 1 : 0;
  return C;
}

void vcycle(Ctx* c, MgHier* H, size_t li)
{
  MgLevel& l = H->L[li];
  StreamScope scope(l.ctx, c->stream);
  const double hi = 1.1 * l.lmax;
  if (li + 1 == H->L.size()) { mg_smooth(c, l, l.b, l.x, true, H->coarse_degree, hi / 30.0, hi); return; }
  MgLevel& n = H->L[li + 1];
  mg_smooth(c, l, l.b, l.x, true, H->degree, hi / 4.0, hi);
  level_spmv(l.ctx, l.ctx->mats[l.mat], l.x, l.t);
  MDB_LAUNCH(c, k_mg_sub, vgrid(l.n), 256, 0, l.n, l.b, l.t, l.r);
  const int dim = l.ctx->mesh.dim;
  const bool part = l.ctx->mesh.partitioned && l.ctx->nranks > 1;
  if (part) halo_exchange(l.ctx, l.r);                   // the restriction to a rank's lowest owned coarse plane reads the fine ghost plane below
  for (size_t k = 0; k < l.ctx->children.size(); ++k) {
    const MgChild C = child_pair(l.ctx, n.ctx, k);
    if (C.size_c > 0 && C.dg) MDB_LAUNCH(c, k_mg_restrict_dg, cdiv(C.size_c, 128), 128, 0, C, dim, l.r, n.b);
    else if (C.size_c > 0) MDB_LAUNCH(c, k_mg_restrict, cdiv(C.size_c, 128), 128, 0, C, dim, l.ctx->d_constrained, n.ctx->d_constrained, l.r, n.b);
  }
  vcycle(c, H, li + 1);
  if (part) { StreamScope cs(n.ctx, c->stream); halo_exchange(n.ctx, n.x); }     // the prolongation reads the coarse ghost planes
  for (size_t k = 0; k < l.ctx->children.size(); ++k) {
    const MgChild C = child_pair(l.ctx, n.ctx, k);
    if (C.size_f > 0 && C.dg) MDB_LAUNCH(c, k_mg_prolong_dg, cdiv(C.size_f, 256), 256, 0, C, dim, n.x, l.x);
    else if (C.size_f > 0) MDB_LAUNCH(c, k_mg_prolong, cdiv(C.size_f, 256), 256, 0, C, dim, l.ctx->d_constrained, n.x, l.x);
  }
  mg_smooth(c, l, l.b, l.x, false, H->degree, hi / 4.0, hi);
}

} // anonymous namespace

void mg_free(Matrix& M)
{
  MgHier* H = (MgHier*)M.mg;
  if (!H) return;
  for (size_t i = 0; i < H->L.size(); ++i) level_free(H->L[i], i == 0);
  if (H->d_red) cudaFree(H->d_red);
  delete H; M.mg = nullptr;
}

static void mg_check(Ctx* c, Matrix& M)
{
  MDB_REQUIRE(!c->um, MDB_EINVAL, "MDB_PRECOND_MG: structured meshes only");
  for (const Child& ch : c->children)
    MDB_REQUIRE((ch.fe_kind == MDB_FE_Q1 && !ch.iface) || (ch.dg && !c->mesh.partitioned), MDB_EINVAL, "MDB_PRECOND_MG: Q1 child spaces only (QkDG on one rank)");
  for (const SubProblemDesc& sp : c->sps)
    MDB_REQUIRE(sp.op_id == MDB_OP_POISSON || sp.op_id == MDB_OP_L2 || sp.op_id == MDB_OP_CONVDIFF_DG, MDB_EINVAL, "MDB_PRECOND_MG: Poisson / L2 / ConvectionDiffusionDG subproblems only");
  for (const CouplingDesc& cp : c->cps) MDB_REQUIRE(cp.op_id == MDB_OP_CVCF_COUPLING || cp.op_id == MDB_OP_PROPFLOW_COUPLING, MDB_EINVAL, "MDB_PRECOND_MG: flow couplings only");
  MDB_REQUIRE(c->con_recorded, MDB_EINVAL, "MDB_PRECOND_MG: needs constraints assembled by mdb_constraints_assemble (the coarse levels replay them)");
  (void)M;
}

// the hierarchy of contexts and matrices (no values yet); collective on a partition
int mg_setup(Ctx* c, Matrix& M)
{
  mg_check(c, M);
  MgHier* H = (MgHier*)M.mg;
  if (H) return (int)H->L.size();
  const int mat_index = (int)(&M - c->mats.data());
  c->phase_begin("mg_setup");
  H = new MgHier(); M.mg = H;
  H->d_red = dalloc<double>(1);
  if (const char* e = getenv("MDB_MG_DEGREE")) H->degree = std::max(1, atoi(e));
  MgLevel top; top.ctx = static_cast<mdb_ctx*>(c); top.owned = false; top.mat = mat_index; top.n = c->ndof;
  level_alloc(top, true);
  H->L.push_back(top);
  for (int lev = 1; lev < 12; ++lev) {
    Ctx* f = H->L.back().ctx;
    mdb_ctx* cc = coarsen(f);
    if (!cc) break;
    MgLevel l; l.ctx = cc; l.owned = true; l.n = cc->ndof;
    int64_t nnz = 0;
    l.mat = mdb_matrix_create(cc, M.gos.data(), (int)M.gos.size(), &nnz);
    if (l.mat < 0) { const std::string msg = mdb_last_error(cc); mdb_destroy(cc); throw Error(l.mat, "multigrid set-up: coarse matrix: " + msg); }
    level_alloc(l, false);
    H->L.push_back(l);
  }
  c->phase_end();
  return (int)H->L.size();
}

mdb_ctx* mg_level_ctx(Matrix& M, int level)
{
  MgHier* H = (MgHier*)M.mg;
  if (!H || level < 0 || level >= (int)H->L.size()) return nullptr;
  return H->L[level].ctx;
}

// z = V-cycle(d): one cycle from a zero initial guess
void mg_apply(Ctx* c, Matrix& M, const double* d_d, double* d_v)
{
  MDB_REQUIRE(!M.asm_log.empty(), MDB_ESTATE, "MDB_PRECOND_MG: assemble the matrix with mdb_jacobian first");
  if (c->nranks > 1) MDB_REQUIRE(M.mg, MDB_ESTATE, "MDB_PRECOND_MG on a partition: call mdb_mg_setup and connect the halo windows of the coarse levels first");
  mg_setup(c, M);
  MgHier* H = (MgHier*)M.mg;
  if (c->nranks > 1)
    for (size_t i = 1; i < H->L.size(); ++i)
      MDB_REQUIRE(H->L[i].ctx->peer_lower || H->L[i].ctx->peer_upper, MDB_ESTATE, "MDB_PRECOND_MG on a partition: the halo windows of a coarse level are not connected (mdb_mg_level + mdb_halo_import)");
  if (H->log != M.asm_log) {
    c->phase_begin("mg_setup");
    H->log = M.asm_log;
    for (size_t i = 0; i < H->L.size(); ++i) level_numeric(c, H, i);
    c->phase_end();
  }
  H->L[0].b = const_cast<double*>(d_d); H->L[0].x = d_v;
  vcycle(c, H, 0);
}

int mg_levels(const Matrix& M) { return M.mg ? (int)((MgHier*)M.mg)->L.size() : 0; }

} // namespace mdb
