// direct.cu — MDB_SOLVER_DIRECT: a sparse direct solve for the systems the Krylov methods of solver.cu do not handle, in place of
// [UPSTREAM] ISTLBackend_SEQ_SuperLU (test/stokesdarcy2.cc:784-788: the Stokes-Darcy system is an indefinite saddle-point matrix with a
// zero pressure block and entries between 1e-10 and 1e7).
//
//   ordering   reverse Cuthill-McKee on the (symmetrised) pattern, host, once per matrix: the mesh is 2-D, so the profile is a band of
//              half width ~ sqrt(n) x (coefficients per cell)
//   storage    LAPACK general-band layout AB[kl + ku + i - j][j] with kl extra rows for the fill of the row interchanges, column major
//   factor     right-looking LU with partial pivoting, one column per step (dgbtf2's algorithm) in ONE persistent cooperative kernel:
//              every block reads column j, finds the pivot and forms the multipliers itself (no broadcast), the trailing columns
//              j+1..ju are dealt out to the blocks (row interchange and rank-1 update fused, a warp-coalesced sweep down the band), then
//              one grid barrier.  The multipliers go to a separate array, so nobody overwrites column j while others still read it.
//              The band window of a step (kl x (kl + ku) doubles) lives in the 126 MB L2.
//   solve      forward / backward substitution with the recorded interchanges in one block each, then one step of iterative
//              refinement against the CSR matrix (residual by the SpMV of solver.cu)
#include "common.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <queue>

namespace cg = cooperative_groups;

namespace mdb {

namespace {

struct Band {
  int64_t n = 0; int kl = 0, ku = 0, ldab = 0;
  int32_t* d_inv = nullptr;      // old index -> position in the band ordering
  double* d_ab = nullptr;        // [n][ldab]
  double* d_l = nullptr;         // [n][kl + 1] multipliers of column j (entry 0: the pivot)
  int32_t* d_ipiv = nullptr;     // row interchanged with row j
  int* d_info = nullptr;
  double* d_w = nullptr;         // permuted right-hand side / solution
  bool factored = false;
};

__global__ void k_band_fill(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ val,
                            const int32_t* __restrict__ inv, int kv, int ldab, double* ab)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t i = inv[r];
  for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
    const int64_t j = inv[colidx[p]];
    ab[j * ldab + kv + i - j] = val[p];
  }
}

// dgbtf2 in one kernel.  Shared: column j (km + 1 doubles) + reduction scratch.
__global__ void __launch_bounds__(512) k_band_factor(int64_t n, int kl, int ku, int ldab, double* ab, double* lbuf, int32_t* ipiv, int* info)
{
  extern __shared__ double sh[];
  double* col = sh;                          // kl + 1
  double* red = sh + kl + 1;                 // 32 values + 32 indices
  int* redi = (int*)(red + 32);
  cg::grid_group grid = cg::this_grid();
  const int kv = kl + ku;
  const int tid = threadIdx.x, nth = blockDim.x;
  int64_t ju = 0;
  for (int64_t j = 0; j < n; ++j) {
    const int km = (int)((n - 1 - j) < kl ? (n - 1 - j) : kl);
    double* cj = ab + j * ldab + kv;
    // column j and its pivot (first entry of largest magnitude, as idamax)
    double best = -1.0; int bi = 0;
    for (int r = tid; r <= km; r += nth) { const double v = cj[r]; col[r] = v; const double a = fabs(v); if (a > best) { best = a; bi = r; } }
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, o); const int oi = __shfl_down_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if ((tid & 31) == 0) { red[tid >> 5] = best; redi[tid >> 5] = bi; }
    __syncthreads();
    if (tid < 32) {
      best = tid < (nth >> 5) ? red[tid] : -1.0; bi = tid < (nth >> 5) ? redi[tid] : 0x7fffffff;
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o); const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (tid == 0) { red[0] = best; redi[0] = bi; }
    }
    __syncthreads();
    const int jp = redi[0];
    const double amax = red[0];
    __syncthreads();
    const int64_t jun = (j + ku + jp) < (n - 1) ? (j + ku + jp) : (n - 1);
    if (jun > ju) ju = jun;
    if (amax == 0.0) {                        // singular column: record it, no elimination (dgbtf2: INFO = j + 1)
      if (blockIdx.x == 0 && tid == 0) { atomicCAS(info, 0, (int)(j + 1)); ipiv[j] = (int32_t)j; lbuf[j * (int64_t)(kl + 1)] = 0.0; }
      grid.sync();
      continue;
    }
    const double pivot = col[jp], head = col[0];
    __syncthreads();
    // multipliers: rows j and j + jp change places, then the subdiagonal is divided by the pivot
    for (int r = 1 + tid; r <= km; r += nth) col[r] = (r == jp ? head : col[r]) / pivot;
    __syncthreads();
    if (blockIdx.x == 0) {
      double* lj = lbuf + j * (int64_t)(kl + 1);
      for (int r = 1 + tid; r <= km; r += nth) lj[r] = col[r];
      if (tid == 0) { lj[0] = pivot; ipiv[j] = (int32_t)(j + jp); }
    }
    // trailing columns j + c, c = 1 .. ju - j: interchange + rank-1 update, one warp per column at a time
    const int ncol = (int)(ju - j);
    const int warps = nth >> 5, wid = tid >> 5, lane = tid & 31;
    for (int c = 1 + blockIdx.x * warps + wid; c <= ncol; c += gridDim.x * warps) {
      double* cc = ab + (j + c) * (int64_t)ldab + kv - c;         // cc[r] = A(j + r, j + c)
      const double t0 = cc[0], tp = cc[jp];
      const double u = tp;                                          // the pivot row's entry of this column
      if (u != 0.0 || jp != 0) {
        for (int r = 1 + lane; r <= km; r += 32) {
          const double v = (r == jp) ? t0 : cc[r];
          cc[r] = v - col[r] * u;
        }
        if (lane == 0) cc[0] = tp;
      }
    }
    grid.sync();
  }
}

// b (permuted) <- L^{-1} P b, then U^{-1}: one block, the column's entries spread over the threads
__global__ void __launch_bounds__(1024) k_band_solve(int64_t n, int kl, int ku, int ldab, const double* __restrict__ ab, const double* __restrict__ lbuf,
                                                     const int32_t* __restrict__ ipiv, double* b)
{
  const int tid = threadIdx.x, nth = blockDim.x;
  const int kv = kl + ku;
  __shared__ double bj;
  for (int64_t j = 0; j < n; ++j) {
    const int km = (int)((n - 1 - j) < kl ? (n - 1 - j) : kl);
    if (tid == 0) { const int64_t p = ipiv[j]; const double t = b[p]; b[p] = b[j]; b[j] = t; bj = t; }
    __syncthreads();
    const double v = bj;
    if (v != 0.0) { const double* lj = lbuf + j * (int64_t)(kl + 1); for (int r = 1 + tid; r <= km; r += nth) b[j + r] -= lj[r] * v; }
    __syncthreads();
  }
  for (int64_t j = n - 1; j >= 0; --j) {
    if (tid == 0) { const double piv = lbuf[j * (int64_t)(kl + 1)]; bj = piv != 0.0 ? b[j] / piv : 0.0; b[j] = bj; }
    __syncthreads();
    const double v = bj;
    const int up = (int)(j < kv ? j : kv);
    if (v != 0.0) { const double* cj = ab + j * (int64_t)ldab + kv; for (int r = 1 + tid; r <= up; r += nth) b[j - r] -= cj[-r] * v; }
    __syncthreads();
  }
}

__global__ void k_permute_in(int64_t n, const int32_t* __restrict__ inv, const double* __restrict__ src, double* dst)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[inv[i]] = src[i];
}
__global__ void k_permute_out(int64_t n, const int32_t* __restrict__ inv, const double* __restrict__ src, double* dst, int add)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) { if (add) dst[i] += src[inv[i]]; else dst[i] = src[inv[i]]; }
}
__global__ void k_axmb(int64_t n, const double* __restrict__ b, double* r)       // r <- b - r
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) r[i] = b[i] - r[i];
}
__global__ void k_norm2(int64_t n, const double* __restrict__ v, double* out)
{
  __shared__ double sh[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += v[i] * v[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i]; out[0] = s; }
}

// reverse Cuthill-McKee of the symmetrised pattern; returns inv[old] = new
std::vector<int32_t> rcm(int64_t n, const std::vector<int64_t>& rowptr, const std::vector<int32_t>& colidx)
{
  std::vector<std::vector<int32_t>> adj(n);
  for (int64_t r = 0; r < n; ++r)
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) { const int32_t cidx = colidx[p]; if (cidx != r) { adj[r].push_back(cidx); adj[cidx].push_back((int32_t)r); } }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  std::vector<int32_t> order; order.reserve(n);
  std::vector<char> seen(n, 0);
  std::vector<int32_t> level(n, -1);
  auto bfs_far = [&](int32_t start, std::vector<int32_t>& comp) {       // BFS; returns a node of minimal degree in the last level
    comp.clear();
    std::vector<int32_t> frontier{start}; level[start] = 0; comp.push_back(start);
    int32_t far = start;
    while (!frontier.empty()) {
      std::vector<int32_t> next;
      for (int32_t v : frontier) for (int32_t w : adj[v]) if (level[w] < 0) { level[w] = level[v] + 1; next.push_back(w); comp.push_back(w); }
      if (!next.empty()) { far = next[0]; for (int32_t w : next) if (adj[w].size() < adj[far].size()) far = w; }
      frontier.swap(next);
    }
    return far;
  };
  std::vector<int32_t> comp;
  for (int64_t s = 0; s < n; ++s) {
    if (seen[s]) continue;
    int32_t start = (int32_t)s;
    for (int it = 0; it < 3; ++it) {                                       // pseudo-peripheral start node
      const int32_t far = bfs_far(start, comp);
      for (int32_t v : comp) level[v] = -1;
      if (far == start) break;
      start = far;
    }
    std::queue<int32_t> q; q.push(start); seen[start] = 1;
    while (!q.empty()) {
      const int32_t v = q.front(); q.pop(); order.push_back(v);
      std::vector<int32_t> nb;
      for (int32_t w : adj[v]) if (!seen[w]) { seen[w] = 1; nb.push_back(w); }
      std::sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) { return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b; });
      for (int32_t w : nb) q.push(w);
    }
  }
  std::vector<int32_t> inv(n);
  for (int64_t i = 0; i < n; ++i) inv[order[n - 1 - i]] = (int32_t)i;
  return inv;
}

void band_free(Band* b)
{
  if (!b) return;
  if (b->d_inv) cudaFree(b->d_inv);
  if (b->d_ab) cudaFree(b->d_ab);
  if (b->d_l) cudaFree(b->d_l);
  if (b->d_ipiv) cudaFree(b->d_ipiv);
  if (b->d_info) cudaFree(b->d_info);
  if (b->d_w) cudaFree(b->d_w);
  delete b;
}

} // anonymous namespace

void direct_free(Matrix& M) { band_free((Band*)M.direct); M.direct = nullptr; }
void direct_invalidate(Matrix& M) { if (M.direct) ((Band*)M.direct)->factored = false; }

int direct_solve(Ctx* c, Matrix& M, const double* d_b, double* d_z, mdb_solve_stats* stats)
{
  MDB_REQUIRE(c->nranks == 1, MDB_EINVAL, "MDB_SOLVER_DIRECT: single-rank contexts only");
  const int64_t n = c->ndof;
  Band* B = (Band*)M.direct;
  cudaEvent_t e0, e1; MDB_CUDA(cudaEventCreate(&e0)); MDB_CUDA(cudaEventCreate(&e1));
  MDB_CUDA(cudaEventRecord(e0, c->stream));
  if (!B) {
    c->phase_begin("direct_order");
    std::vector<int64_t> rowptr(n + 1); std::vector<int32_t> colidx(M.nnz);
    MDB_CUDA(cudaMemcpyAsync(rowptr.data(), M.d_rowptr, (n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(colidx.data(), M.d_colidx, M.nnz * 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    const std::vector<int32_t> inv = rcm(n, rowptr, colidx);
    int64_t bw = 0;
    for (int64_t r = 0; r < n; ++r) for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) bw = std::max<int64_t>(bw, std::llabs((long long)inv[r] - inv[colidx[p]]));
    B = new Band(); M.direct = B;
    B->n = n; B->kl = B->ku = (int)bw; B->ldab = 2 * B->kl + B->ku + 1;
    const size_t bytes = (size_t)n * B->ldab * 8 + (size_t)n * (B->kl + 1) * 8;
    size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
    MDB_REQUIRE(bytes < free_b * 0.9, MDB_ENOMEM, "MDB_SOLVER_DIRECT: the band of the reordered matrix (half width " + std::to_string(bw) + ") does not fit the device memory");
    B->d_inv = dalloc<int32_t>(n); B->d_ab = dalloc<double>((size_t)n * B->ldab); B->d_l = dalloc<double>((size_t)n * (B->kl + 1));
    B->d_ipiv = dalloc<int32_t>(n); B->d_info = dalloc<int>(1); B->d_w = dalloc<double>(2 * n);
    MDB_CUDA(cudaMemcpyAsync(B->d_inv, inv.data(), n * 4, cudaMemcpyHostToDevice, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    c->phase_end();
  }
  if (!B->factored) {
    c->phase_begin("direct_factor");
    MDB_CUDA(cudaMemsetAsync(B->d_ab, 0, (size_t)n * B->ldab * 8, c->stream));
    MDB_CUDA(cudaMemsetAsync(B->d_info, 0, sizeof(int), c->stream));
    MDB_LAUNCH(c, k_band_fill, cdiv(n, 128), 128, 0, n, M.d_rowptr, M.d_colidx, M.d_val, B->d_inv, B->kl + B->ku, B->ldab, B->d_ab);
    const size_t smem = (size_t)(B->kl + 1) * 8 + 32 * 8 + 32 * 4;
    MDB_REQUIRE(smem <= 200 * 1024, MDB_EINVAL, "MDB_SOLVER_DIRECT: band too wide for the factorisation kernel");
    MDB_CUDA(cudaFuncSetAttribute(k_band_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0; cudaGetDevice(&dev);
    MDB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_band_factor, 512, smem));
    MDB_REQUIRE(per_sm >= 1, MDB_EINVAL, "MDB_SOLVER_DIRECT: the factorisation kernel does not fit an SM");
    const int grid = sms * (per_sm > 2 ? 2 : per_sm);
    int64_t nn = n; int kl = B->kl, ku = B->ku, ldab = B->ldab;
    void* args[] = {&nn, &kl, &ku, &ldab, &B->d_ab, &B->d_l, &B->d_ipiv, &B->d_info};
    MDB_CUDA(cudaLaunchCooperativeKernel((void*)k_band_factor, dim3(grid), dim3(512), args, smem, c->stream)); c->launches++;
    c->phase_end();
    B->factored = true;
  }
  c->phase_begin("direct_solve");
  double* w = B->d_w; double* res = B->d_w + n;
  MDB_LAUNCH(c, k_norm2, 1, 1024, 0, n, d_b, c->d_scal + 48);
  MDB_LAUNCH(c, k_permute_in, cdiv(n, 256), 256, 0, n, B->d_inv, d_b, w);
  MDB_LAUNCH(c, k_band_solve, 1, 1024, 0, n, B->kl, B->ku, B->ldab, B->d_ab, B->d_l, B->d_ipiv, w);
  MDB_LAUNCH(c, k_permute_out, cdiv(n, 256), 256, 0, n, B->d_inv, w, d_z, 0);
  for (int it = 0; it < 2; ++it) {                          // iterative refinement; the last pass only measures the defect
    spmv(c, M, d_z, res);
    MDB_LAUNCH(c, k_axmb, cdiv(n, 256), 256, 0, n, d_b, res);
    MDB_LAUNCH(c, k_norm2, 1, 1024, 0, n, res, c->d_scal + 49 + it);
    if (it == 1) break;
    MDB_LAUNCH(c, k_permute_in, cdiv(n, 256), 256, 0, n, B->d_inv, res, w);
    MDB_LAUNCH(c, k_band_solve, 1, 1024, 0, n, B->kl, B->ku, B->ldab, B->d_ab, B->d_l, B->d_ipiv, w);
    MDB_LAUNCH(c, k_permute_out, cdiv(n, 256), 256, 0, n, B->d_inv, w, d_z, 1);
  }
  c->phase_end();
  MDB_CUDA(cudaEventRecord(e1, c->stream));
  int info = 0;
  MDB_CUDA(cudaMemcpyAsync(&info, B->d_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(c->h_pinned + 48, c->d_scal + 48, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (stats) {
    stats->iterations = 1; stats->defect0 = std::sqrt(c->h_pinned[48]); stats->defect = std::sqrt(c->h_pinned[50]);
    stats->converged = (info == 0 && std::isfinite(stats->defect)) ? 1 : 0; stats->seconds = ms * 1e-3;
  }
  if (info != 0) { c->err = "MDB_SOLVER_DIRECT: the matrix is singular (zero pivot column " + std::to_string(info - 1) + " of the band ordering)"; return MDB_ENOTCONVERGED; }
  return MDB_OK;
}

} // namespace mdb
