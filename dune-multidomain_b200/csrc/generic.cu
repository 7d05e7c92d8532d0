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
#include <cub/cub.cuh>
#include <vector>
#include <algorithm>

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

// ---- simple rows of a Qk child: one representative per class, replicated ---------------------------------------------------------
// On a structured mesh the rows of a Qk child fall into k^d classes by the position of their lattice point inside the cell (vertex,
// edge along x, ..., cell interior).  A row is SIMPLE when every cell around its lattice point exists and carries the child's
// subdomain byte, and the row holds nothing but its own block's volume links (length = prod_d (a_d == 0 ? 2k + 1 : k + 1)): all simple
// rows of a class then hold the same values in the same column order (the numbering is lexicographic inside the entity groups), and
// consecutive simple rows of a class form long runs of the value array (the DOFs of a group are numbered consecutively).  Per call
// the owner-computes kernel evaluates the irregular rows in place and ONE representative row per class into a scratch image; a
// streaming kernel replicates the images over the runs with aligned 256-byte warp stores — the HBM-write-bound path of the Q1 rows
// (k_jacobian_fill_q1) for every Qk.  Built once per (matrix, child).
#define QK_MAXCLS 8
#define QK_MAXL 128
struct QkRun { long long p0, p1; int cls, ns; };
struct QkPlan {
  int ncls = 0; int64_t nruns = 0, nirr = 0, nrep = 0;
  int32_t* d_irr = nullptr;          // irregular rows (child-local index), ascending
  int32_t* d_rep = nullptr;          // representative rows per class (child-local index; -1: the class has no simple row)
  QkRun* d_runs = nullptr;
  double* d_img = nullptr;           // QK_MAXCLS x QK_MAXL scratch images
  bool usable = false;
};
void qk_plan_free(void* p)
{
  QkPlan* P = (QkPlan*)p;
  if (!P) return;
  if (P->d_irr) cudaFree(P->d_irr); if (P->d_rep) cudaFree(P->d_rep); if (P->d_runs) cudaFree(P->d_runs); if (P->d_img) cudaFree(P->d_img);
  delete P;
}

// cls[gi] = class of a simple row, 255 otherwise; rep[class] = smallest simple row of the class
template <int DIM>
__global__ void k_qk_classify(MeshDesc m, const uint8_t* __restrict__ sd, ChildG ch, uint8_t sd0, const int64_t* __restrict__ rowptr, uint8_t* cls, int* rep)
{
  const int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gi >= ch.size) return;
  const int k = ch.k;
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  int klass = 0, mul = 1, len = 1;
  bool ok = true;
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (d >= DIM) continue;
    const int a = p[d] % k, c1 = p[d] / k;
    klass += a * mul; mul *= k;
    if (a == 0) { lo[d] = c1 - 1; hi[d] = c1; len *= 2 * k + 1; if (c1 - 1 < 0 || c1 >= m.n[d]) ok = false; }
    else { lo[d] = hi[d] = c1; len *= k + 1; }
  }
  if (ok)
    for (int cz = lo[2]; cz <= hi[2]; ++cz)
      for (int cy = lo[1]; cy <= hi[1]; ++cy)
        for (int cx = lo[0]; cx <= hi[0]; ++cx)
          ok = ok && sd[cx + (int64_t)m.n[0] * (cy + (int64_t)m.n[1] * cz)] == sd0;
  const int64_t g = ch.offset + gi;
  ok = ok && (rowptr[g + 1] - rowptr[g] == len) && len <= QK_MAXL && klass < QK_MAXCLS;
  cls[gi] = ok ? (uint8_t)klass : (uint8_t)255;
  if (ok) atomicMin(rep + klass, (int)gi);
}
// first cell of the child's subdomain -> its subdomain byte
__global__ void k_qk_first_sd(const uint8_t* __restrict__ sd, int64_t ncells, int subdomain, unsigned long long* out)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ncells; i += (int64_t)gridDim.x * blockDim.x)
    if (subdomain < 0 || ((sd[i] >> subdomain) & 1)) { atomicMin(out, ((unsigned long long)i << 8) | sd[i]); return; }      // ascending per thread: the first hit is its smallest
}
__global__ void k_qk_flags(const uint8_t* __restrict__ cls, int64_t n, int32_t* irr, int32_t* head, int32_t* tail)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  irr[i] = cls[i] == 255 ? 1 : 0;
  head[i] = (cls[i] != 255 && (i == 0 || cls[i - 1] != cls[i])) ? 1 : 0;          // a run of simple rows of one class starts here
  tail[i] = (cls[i] != 255 && (i == n - 1 || cls[i + 1] != cls[i])) ? 1 : 0;      // ... and ends here
}
__global__ void k_qk_compact(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int64_t n, int32_t* list)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && flags[i]) list[scan[i]] = (int32_t)i;
}
__global__ void k_qk_runs(const int32_t* __restrict__ start, const int32_t* __restrict__ last, int64_t nruns, const uint8_t* __restrict__ cls, int64_t child_offset,
                          const int64_t* __restrict__ rowptr, QkRun* runs)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nruns) return;
  const int64_t a = start[i], b = (int64_t)last[i] + 1;
  QkRun R; R.p0 = rowptr[child_offset + a]; R.p1 = rowptr[child_offset + b]; R.cls = cls[a];
  R.ns = (int)(rowptr[child_offset + a + 1] - rowptr[child_offset + a]);
  runs[i] = R;
}

// images[cls] -> every run of the class.  Same store pattern as k_jacobian_fill_q1: chunks of 32 rows' worth of doubles on 256-byte
// boundaries, the two edges of a run masked; the row length is a run-time value here.
__global__ void __launch_bounds__(128, 12) k_qk_fill(const QkRun* __restrict__ runs, int64_t nruns, const double* __restrict__ images, double* __restrict__ val, int overwrite)
{
  __shared__ double img[QK_MAXCLS * QK_MAXL];
  for (int i = threadIdx.x; i < QK_MAXCLS * QK_MAXL; i += blockDim.x) img[i] = images[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t run = blockIdx.x; run < nruns; run += gridDim.x) {
    const QkRun R = runs[run];
    const int ns = R.ns;
    const double* S = img + R.cls * QK_MAXL;
    const int chunk = 32 * ns;
    const int64_t p0 = R.p0, p1 = R.p1, a0 = p0 & ~(int64_t)31;
    const int inc = 32 % ns;
    const int mm0 = (int)((a0 - p0 + lane + 32 * (int64_t)ns) % ns);
    for (int64_t cb = a0 + (int64_t)warp * chunk; cb < p1; cb += (int64_t)(blockDim.x >> 5) * chunk) {
      int mm = mm0;
      if (cb >= p0 && cb + chunk <= p1) {
        if (overwrite) for (int it = 0; it < ns; ++it) { val[cb + it * 32 + lane] = S[mm]; mm += inc; if (mm >= ns) mm -= ns; }
        else for (int it = 0; it < ns; ++it) { val[cb + it * 32 + lane] += S[mm]; mm += inc; if (mm >= ns) mm -= ns; }
      } else {
        for (int it = 0; it < ns; ++it) {
          const int64_t i = cb + it * 32 + lane;
          if (i >= p0 && i < p1) { if (overwrite) val[i] = S[mm]; else val[i] += S[mm]; }
          mm += inc; if (mm >= ns) mm -= ns;
        }
      }
    }
  }
}

// rows of a list (child-local indices); scratch != null: row i of the list goes to scratch + i * QK_MAXL with overwrite semantics.
// One WARP per row: lane lj handles local coefficient lj of every incident cell (the columns of one cell are distinct, so the lanes never
// meet), the cells are visited one after the other in ascending order — every entry is summed in the reference's order, and the
// dependent chain of a row is (cells) x (one binary search) instead of (cells) x (local coefficients) x (one binary search).
template <int DIM>
__global__ void __launch_bounds__(128) k_jacobian_rows_qk_list(MeshDesc m, const uint8_t* __restrict__ sd, ChildG ch, VolSpecG V, const double* __restrict__ Ae,
                                                               const int32_t* __restrict__ list, int64_t nlist, const int64_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ colidx, double* __restrict__ val, int overwrite, double* __restrict__ scratch)
{
  const int64_t li_ = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (li_ >= nlist) return;
  const int64_t gi = list[li_];
  if (gi < 0) return;
  const int64_t g = ch.offset + gi;
  const int64_t rs = rowptr[g], re = rowptr[g + 1];
  double* out = scratch ? scratch + li_ * QK_MAXL - rs : val;
  if (overwrite || scratch) for (int64_t q = rs + lane; q < re; q += 32) out[q] = 0.0;
  __syncwarp();
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  for_incident_cells<DIM>(m, sd, ch, p, [&](const int* c, uint8_t s, int li) {
    for (int lj = lane; lj < ch.nloc; lj += 32) {
      const int32_t col = cell_dof<DIM>(ch, c, lj);
      int64_t lo = rs, hi = re;
      while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (colidx[mid] < col) lo = mid + 1; else hi = mid; }
      for (int v = 0; v < V.n; ++v) {
        if (!cond_applies(V.cond_kind[v], V.mask[v], s)) continue;
        out[lo] += Ae[(size_t)V.slot[v] * MDB_MAXLOC * MDB_MAXLOC + (size_t)li * ch.nloc + lj];
      }
    }
    __syncwarp();                                                             // the next cell's lanes may hit the columns this cell's lanes wrote
  });
}

static QkPlan* qk_plan(Ctx* c, Matrix& M, int child, const ChildG& ch)
{
  if (M.qk_plans[child]) return (QkPlan*)M.qk_plans[child];
  const MeshDesc& m = c->mesh;
  QkPlan* P = new QkPlan(); M.qk_plans[child] = P;
  int ncls = 1; for (int d = 0; d < m.dim; ++d) ncls *= ch.k;
  if (ncls > QK_MAXCLS || ch.size >= (int64_t)2147483647) return P;
  P->ncls = ncls;
  const int64_t n = ch.size;
  unsigned long long* d_first = dalloc<unsigned long long>(1);
  MDB_CUDA(cudaMemsetAsync(d_first, 0xFF, sizeof(unsigned long long), c->stream));
  MDB_LAUNCH(c, k_qk_first_sd, 148, 256, 0, c->d_cell_sd, m.ncells(), ch.subdomain, d_first);
  unsigned long long first = 0;
  MDB_CUDA(cudaMemcpyAsync(&first, d_first, sizeof(first), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  dfree(d_first);
  if (first == ~0ull) return P;
  const uint8_t sd0 = (uint8_t)(first & 255u);
  uint8_t* d_cls = dalloc<uint8_t>(n);
  P->d_rep = dalloc<int32_t>(QK_MAXCLS);
  MDB_CUDA(cudaMemsetAsync(P->d_rep, 0x7F, QK_MAXCLS * sizeof(int32_t), c->stream));
  if (m.dim == 2) MDB_LAUNCH(c, k_qk_classify<2>, cdiv(n, 256), 256, 0, m, c->d_cell_sd, ch, sd0, M.d_rowptr, d_cls, P->d_rep);
  else MDB_LAUNCH(c, k_qk_classify<3>, cdiv(n, 256), 256, 0, m, c->d_cell_sd, ch, sd0, M.d_rowptr, d_cls, P->d_rep);
  int32_t rep[QK_MAXCLS];
  MDB_CUDA(cudaMemcpyAsync(rep, P->d_rep, sizeof(rep), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < QK_MAXCLS; ++k) if (rep[k] == 0x7F7F7F7F) rep[k] = -1;
  MDB_CUDA(cudaMemcpyAsync(P->d_rep, rep, sizeof(rep), cudaMemcpyHostToDevice, c->stream));
  int32_t* d_irr = dalloc<int32_t>(n); int32_t* d_head = dalloc<int32_t>(n); int32_t* d_tail = dalloc<int32_t>(n); int32_t* d_scan = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_qk_flags, cdiv(n, 256), 256, 0, d_cls, n, d_irr, d_head, d_tail);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_irr, d_scan, (int)n, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  auto count = [&](int32_t* flags) {
    MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, flags, d_scan, (int)n, c->stream)); c->launches++;
    int32_t ls = 0, lf = 0;
    MDB_CUDA(cudaMemcpyAsync(&ls, d_scan + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(&lf, flags + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    return (int64_t)ls + lf;
  };
  P->nirr = count(d_irr);
  P->d_irr = dalloc<int32_t>(P->nirr > 0 ? P->nirr : 1);
  MDB_LAUNCH(c, k_qk_compact, cdiv(n, 256), 256, 0, d_irr, d_scan, n, P->d_irr);
  P->nruns = count(d_head);
  int32_t* d_start = dalloc<int32_t>(P->nruns > 0 ? P->nruns : 1);
  MDB_LAUNCH(c, k_qk_compact, cdiv(n, 256), 256, 0, d_head, d_scan, n, d_start);
  const int64_t nlast = count(d_tail);
  MDB_REQUIRE(nlast == P->nruns, MDB_EINVAL, "Qk plan: run heads and tails do not pair up");
  int32_t* d_last = dalloc<int32_t>(P->nruns > 0 ? P->nruns : 1);
  MDB_LAUNCH(c, k_qk_compact, cdiv(n, 256), 256, 0, d_tail, d_scan, n, d_last);
  P->d_runs = dalloc<QkRun>(P->nruns > 0 ? P->nruns : 1);
  if (P->nruns > 0) MDB_LAUNCH(c, k_qk_runs, cdiv(P->nruns, 128), 128, 0, d_start, d_last, P->nruns, d_cls, ch.offset, M.d_rowptr, P->d_runs);
  // long runs (a whole entity group can be one run) are cut into pieces of whole rows so that the streaming kernel's blocks share them
  if (P->nruns > 0) {
    std::vector<QkRun> h(P->nruns), pieces;
    MDB_CUDA(cudaMemcpyAsync(h.data(), P->d_runs, P->nruns * sizeof(QkRun), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    for (const QkRun& R : h) {
      const long long piece = (long long)std::max(1, 32768 / R.ns) * R.ns;
      for (long long q = R.p0; q < R.p1; q += piece) { QkRun r = R; r.p0 = q; r.p1 = std::min(q + piece, R.p1); pieces.push_back(r); }
    }
    cudaFree(P->d_runs);
    P->nruns = (int64_t)pieces.size();
    P->d_runs = dalloc<QkRun>(P->nruns);
    MDB_CUDA(cudaMemcpyAsync(P->d_runs, pieces.data(), pieces.size() * sizeof(QkRun), cudaMemcpyHostToDevice, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
  }
  P->d_img = dalloc<double>(QK_MAXCLS * QK_MAXL);
  MDB_CUDA(cudaMemsetAsync(P->d_img, 0, QK_MAXCLS * QK_MAXL * sizeof(double), c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_t); dfree(d_cls); dfree(d_irr); dfree(d_head); dfree(d_tail); dfree(d_scan); dfree(d_start); dfree(d_last);
  P->usable = P->nruns > 0;
  if (getenv("MDB_QK_DEBUG")) fprintf(stderr, "qk plan child %d: rows %lld irregular %lld runs %lld classes %d sd0 %d reps %d %d %d %d\n", child, (long long)n, (long long)P->nirr,
                                      (long long)P->nruns, ncls, (int)sd0, rep[0], rep[1], rep[2], rep[3]);
  return P;
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
  static const bool no_plan = getenv("MDB_QK_ALL_ROWS") != nullptr;               // A/B switch: every row through the owner-computes kernel
  QkPlan* P = no_plan ? nullptr : qk_plan(c, M, child, ch);
  if (P && P->usable) {
    if (P->nirr > 0) {
      if (m.dim == 2) MDB_LAUNCH(c, k_jacobian_rows_qk_list<2>, cdiv(P->nirr, 4), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, P->d_irr, P->nirr, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, (double*)nullptr);
      else MDB_LAUNCH(c, k_jacobian_rows_qk_list<3>, cdiv(P->nirr, 4), 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, P->d_irr, P->nirr, M.d_rowptr, M.d_colidx, M.d_val, (int)overwrite, (double*)nullptr);
    }
    if (m.dim == 2) MDB_LAUNCH(c, k_jacobian_rows_qk_list<2>, QK_MAXCLS / 4, 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, P->d_rep, (int64_t)QK_MAXCLS, M.d_rowptr, M.d_colidx, M.d_val, 1, P->d_img);
    else MDB_LAUNCH(c, k_jacobian_rows_qk_list<3>, QK_MAXCLS / 4, 128, 0, m, c->d_cell_sd, ch, V, c->d_Ae, P->d_rep, (int64_t)QK_MAXCLS, M.d_rowptr, M.d_colidx, M.d_val, 1, P->d_img);
    int grid = 148 * 12; if (grid > P->nruns) grid = (int)P->nruns;
    MDB_LAUNCH(c, k_qk_fill, grid, 128, 0, P->d_runs, P->nruns, P->d_img, M.d_val, (int)overwrite);
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
