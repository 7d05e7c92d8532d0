// solver.cu — Krylov solve of the assembled system: CSR SpMV, fused vector kernels, CG and BiCGSTAB with Jacobi.
//
// Replaces [UPSTREAM] dune-istl CGSolver / BiCGSTABSolver + SeqJac behind PDELab's ISTLBackend_SEQ_* as called from
// StationaryLinearProblemSolver (test/testpoisson.cc:291-298, test/explicitlycoupledpoisson.cc:297-313).  The iteration
// formulas and their order follow ISTL (same scalars, same update order); reductions are two-stage and deterministic
// (block partials in a fixed order), all Krylov scalars stay on the device, and the host only polls a "done" flag every
// few iterations, so the loop is a pure stream of kernels.
//
// SpMV is "CSR-stream": a block streams the contiguous val/col span of its rows fully coalesced, multiplies by the
// gathered x (L1/L2 hits: neighbouring rows share columns) into shared memory, then one thread per row sums its segment.
// Algorithmic bytes: 12*nnz + 20*nrows.
#include "common.cuh"
#include <cstring>
#include <cmath>
#include <cub/cub.cuh>

namespace mdb {

#define SPMV_ROWS 64
#define SPMV_THREADS 256
#define VEC_THREADS 256
#define MAX_PARTIAL_BLOCKS 2048

// scalar slots
// slots reduced together are adjacent (RHONEW,RR), (RR,PQ), (TR,TT): one ncclAllReduce serves a pair
enum { S_RHO = 0, S_RHOLAST, S_RHONEW, S_RR, S_PQ, S_RR0, S_BETA, S_LAMBDA, S_DONE, S_ITER, S_RED2, S_ALPHA, S_OMEGA, S_H, S_TR, S_TT, S_HALF, S_BREAK };

__device__ __forceinline__ double block_sum(double v, double* sh)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
  return s;     // valid in thread 0
}

// streaming loads: val/col are touched exactly once per SpMV, keep them out of L1 so that it serves the x gathers
__device__ __forceinline__ double ld_stream_f64(const double* p)
{
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int32_t* p)
{
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---- variant A: CSR-stream --------------------------------------------------------------------------------
// A block owns SPMV_ROWS consecutive rows.  Phase 1: all threads stream the contiguous val/col span of those rows
// (perfectly coalesced, SPMV_UNROLL independent loads in flight per thread), multiply with the gathered x and park the
// products in shared memory.  Phase 2: SPMV_LPR lanes per row sum the row's segment and combine by shuffles.
// y = A x  (+ optional fused dots: partial[0][b] = sum y_i a_i ; partial[1][b] = sum y_i y_i over owned rows)
#define SPMV_LPR (SPMV_THREADS / SPMV_ROWS)
#define SPMV_UNROLL 8
template <int DOTS>
__global__ void __launch_bounds__(SPMV_THREADS) k_spmv(int64_t n, int nchunks, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                        const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                        const double* __restrict__ a, const uint8_t* __restrict__ owned, double* partial,
                                                        const double* __restrict__ done)
{
  extern __shared__ double sprod[];
  __shared__ double sh[SPMV_THREADS / 32];
  if (done && done[S_DONE] != 0.0) return;
  double d0 = 0.0, d1 = 0.0;
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t r0 = (int64_t)chunk * SPMV_ROWS;
    const int64_t r1 = (r0 + SPMV_ROWS < n) ? r0 + SPMV_ROWS : n;
    const int64_t span0 = rowptr[r0];
    const int span = (int)(rowptr[r1] - span0);
    __syncthreads();
    for (int base = 0; base < span; base += SPMV_THREADS * SPMV_UNROLL) {
      int col[SPMV_UNROLL]; double v[SPMV_UNROLL];
#pragma unroll
      for (int u = 0; u < SPMV_UNROLL; ++u) {
        const int i = base + u * SPMV_THREADS + threadIdx.x;
        col[u] = (i < span) ? ld_stream_s32(colidx + span0 + i) : -1;
        v[u] = (i < span) ? ld_stream_f64(val + span0 + i) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < SPMV_UNROLL; ++u) {
        const int i = base + u * SPMV_THREADS + threadIdx.x;
        if (col[u] >= 0) sprod[i] = v[u] * x[col[u]];
      }
    }
    __syncthreads();
    const int64_t r = r0 + threadIdx.x / SPMV_LPR;
    const int sub = threadIdx.x % SPMV_LPR;
    double sum = 0.0;
    if (r < r1) {
      const int s = (int)(rowptr[r] - span0), e = (int)(rowptr[r + 1] - span0);
      for (int i = s + sub; i < e; i += SPMV_LPR) sum += sprod[i];
    }
#pragma unroll
    for (int o = SPMV_LPR / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, SPMV_LPR);
    if (r < r1 && sub == 0) {
      y[r] = sum;
      if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
    }
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + blockIdx.x] = s1;
    }
  }
}

// ---- variant B: vector CSR, LPR lanes per row, no shared memory (all of L1 serves the x gathers) -----------
template <int DOTS, int LPR>
__global__ void __launch_bounds__(256) k_spmv_vec(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                  const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                  const double* __restrict__ a, const uint8_t* __restrict__ owned, double* partial,
                                                  const double* __restrict__ done)
{
  __shared__ double sh[8];
  if (done && done[S_DONE] != 0.0) return;
  double d0 = 0.0, d1 = 0.0;
  const int sub = threadIdx.x % LPR;
  const int64_t rows_per_grid = (int64_t)gridDim.x * (256 / LPR);
  for (int64_t rb = blockIdx.x * (int64_t)(256 / LPR); rb < n; rb += rows_per_grid) {     // block-uniform trip count
    const int64_t r = rb + threadIdx.x / LPR;
    double sum = 0.0;
    if (r < n) {
      const int64_t s = rowptr[r], e = rowptr[r + 1];
      for (int64_t p = s + sub; p < e; p += 4 * LPR) {
        int c0 = ld_stream_s32(colidx + p);
        int c1 = (p + LPR < e) ? ld_stream_s32(colidx + p + LPR) : -1;
        int c2 = (p + 2 * LPR < e) ? ld_stream_s32(colidx + p + 2 * LPR) : -1;
        int c3 = (p + 3 * LPR < e) ? ld_stream_s32(colidx + p + 3 * LPR) : -1;
        double v0 = ld_stream_f64(val + p);
        double v1 = (c1 >= 0) ? ld_stream_f64(val + p + LPR) : 0.0;
        double v2 = (c2 >= 0) ? ld_stream_f64(val + p + 2 * LPR) : 0.0;
        double v3 = (c3 >= 0) ? ld_stream_f64(val + p + 3 * LPR) : 0.0;
        sum += v0 * x[c0];
        if (c1 >= 0) sum += v1 * x[c1];
        if (c2 >= 0) sum += v2 * x[c2];
        if (c3 >= 0) sum += v3 * x[c3];
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, LPR);
    if (r < n && sub == 0) {
      y[r] = sum;
      if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
    }
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + blockIdx.x] = s1;
    }
  }
}

// ---- variant C: TMA-fed, thread-per-row, stencil-aware (the default on Q1 spaces) --------------------------------------
// A block owns ST_ROWS consecutive rows of one child block = ONE contiguous span of the value array.  One elected thread
// fetches that span AND the 3^(d-1) segments of x the rows read into shared memory with bulk-async copies (cp.async.bulk, the
// 1-D TMA path) on one mbarrier: no per-thread global loads on the hot path, the bytes arrive while the other resident blocks
// compute.  Every thread then owns one row and sums it sequentially in column order (= the summation order of a
// sequential CSR product).
// Why the x segments are known without colidx: when the DOF-carrying lattice points of a child block fill their bounding box
// (half-space subdomains: always), the block is numbered lexicographically over that box
// (multidomaingridfunctionspace.hh:308-369), so a row r whose columns all lie in its own block ("affine": every row but the
// interface rows) has the columns r + dx + dy*bx + dz*bx*by for the set bits (dx,dy,dz) of its stencil mask (rowinfo), in
// ascending order.  The 108 bytes of colidx per row are never read, and ST_ROWS consecutive rows read 3^(d-1) contiguous
// segments of x of ST_ROWS + 2 entries.  The other rows (~1.5 %) walk their colidx entries, 8 lanes per row.
// Traffic per simple row: 216 B values + 12 B (rowptr, rowinfo) + 16 B vectors instead of 216 + 108 + 28.
#define ST_ROWS 128
#define ST_BLOCKS 5
#define ST_VCAP (ST_ROWS * 28 + 2)           // doubles of value staging per block (simple rows: 27 per row + alignment slack)
#define ST_XSEG (ST_ROWS + 4)                // doubles per staged x segment
// chunk0: first VIRTUAL chunk of the child in this launch.  A launch covers all chunks of the child (skip_len = 0), only its interior
// chunks [a, b) (skip_lo = 0, skip_len = a, b - a virtual chunks) or all but those (skip_lo = a, skip_len = b - a): virtual chunk v of
// the child is its chunk v < skip_lo ? v : v + skip_len.  On a partition the interior chunks read no ghost entry of x, so they run
// while the halo planes are in flight (solver: push -> interior -> wait -> the rest).
struct SpmvChild { int64_t offset, size, chunk0, skip_lo, skip_len; int box, bx, bxy; };
struct SpmvSpace { int nchild; int64_t nchunks; SpmvChild ch[MDB_MAX_CHILDREN]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// x segments of the chunk [r0, r1) of child ch: entries [r0 - 1 + off, r1 + 1 + off) for off = dy*bx + dz*bxy.  The chunk takes
// the staged path only when every segment (widened by one entry for 16-byte alignment) lies inside the vector.
__host__ __device__ inline bool spmv_chunk_staged(const SpmvChild& ch, int dim, int64_t r0, int64_t r1, int64_t n)
{
  const int64_t reach = (dim == 3) ? (int64_t)ch.bx + ch.bxy : ch.bx;
  return r1 > r0 && ch.box && r0 - 1 - reach - 1 >= 0 && r1 + 1 + reach + 1 <= n;
}
// a row the stencil kernel sums: its chunk is staged and all its columns are own-block stencil neighbours
__device__ __forceinline__ uint32_t spmv_affine_mask(int dim, uint32_t info, int64_t rowlen)
{
  const uint32_t m = (dim == 3) ? (info & 0x7FFFFFFu) : ((info >> 9) & 0x1FFu);
  return rowlen == __popc(m) ? m : 0u;
}

// flags[r] = 1 for every row the stencil kernel does NOT sum
__global__ void k_spmv_rest_flags(int64_t n, SpmvSpace sp, int dim, const int64_t* __restrict__ rowptr, const uint32_t* __restrict__ rowinfo, int32_t* flags)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int ci = 0;
  while (ci + 1 < sp.nchild && r >= sp.ch[ci + 1].offset) ++ci;
  const SpmvChild& ch = sp.ch[ci];
  const int64_t r0 = ch.offset + (r - ch.offset) / ST_ROWS * ST_ROWS;
  const int64_t r1 = (r0 + ST_ROWS < ch.offset + ch.size) ? r0 + ST_ROWS : ch.offset + ch.size;
  const bool mine = spmv_chunk_staged(ch, dim, r0, r1, n) && spmv_affine_mask(dim, rowinfo[r], rowptr[r + 1] - rowptr[r]) != 0u;
  flags[r] = mine ? 0 : 1;
}
__global__ void k_spmv_rest_compact(int64_t n, const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int32_t* list)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n && flags[r]) list[scan[r]] = (int32_t)r;
}

// y[r] = (A x)[r] for the rows of a list, 8 lanes per row (the vector-CSR kernel over a row list); partial sums of the fused
// dots go to partial[pofs + block]
template <int DOTS>
__global__ void __launch_bounds__(256) k_spmv_rows(int64_t nlist, const int32_t* __restrict__ list, const int64_t* __restrict__ rowptr,
                                                   const int32_t* __restrict__ colidx, const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ y, const double* __restrict__ a, const uint8_t* __restrict__ owned, double* partial,
                                                   int pofs, const double* __restrict__ done)
{
  constexpr int LPR = 8;
  __shared__ double sh[8];
  if (done && done[S_DONE] != 0.0) return;
  double d0 = 0.0, d1 = 0.0;
  const int sub = threadIdx.x % LPR;
  const int64_t rows_per_grid = (int64_t)gridDim.x * (256 / LPR);
  for (int64_t ib = blockIdx.x * (int64_t)(256 / LPR); ib < nlist; ib += rows_per_grid) {     // block-uniform trip count
    const int64_t i = ib + threadIdx.x / LPR;
    double sum = 0.0;
    int64_t r = -1;
    if (i < nlist) {
      r = list[i];
      const int64_t s = rowptr[r], e = rowptr[r + 1];
      for (int64_t p = s + sub; p < e; p += 4 * LPR) {
        int c0 = ld_stream_s32(colidx + p);
        int c1 = (p + LPR < e) ? ld_stream_s32(colidx + p + LPR) : -1;
        int c2 = (p + 2 * LPR < e) ? ld_stream_s32(colidx + p + 2 * LPR) : -1;
        int c3 = (p + 3 * LPR < e) ? ld_stream_s32(colidx + p + 3 * LPR) : -1;
        double v0 = ld_stream_f64(val + p);
        double v1 = (c1 >= 0) ? ld_stream_f64(val + p + LPR) : 0.0;
        double v2 = (c2 >= 0) ? ld_stream_f64(val + p + 2 * LPR) : 0.0;
        double v3 = (c3 >= 0) ? ld_stream_f64(val + p + 3 * LPR) : 0.0;
        sum += v0 * x[c0];
        if (c1 >= 0) sum += v1 * x[c1];
        if (c2 >= 0) sum += v2 * x[c2];
        if (c3 >= 0) sum += v3 * x[c3];
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, LPR);
    if (r >= 0 && sub == 0) {
      y[r] = sum;
      if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
    }
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[pofs + blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + pofs + blockIdx.x] = s1;
    }
  }
}

template <int DIM, int DOTS>
__global__ void __launch_bounds__(ST_ROWS, ST_BLOCKS) k_spmv_stencil(int64_t n, SpmvSpace sp, const int64_t* __restrict__ rowptr,
                                                                     const uint32_t* __restrict__ rowinfo, const double* __restrict__ val,
                                                                     const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ a,
                                                                     const uint8_t* __restrict__ owned, double* partial, int pofs,
                                                                     const double* __restrict__ done)
{
  constexpr int NS = DIM == 3 ? 27 : 9;
  constexpr int NL = NS / 3;                                                      // x segments: one per (dy, dz)
  __shared__ __align__(16) double sv[ST_VCAP];
  __shared__ __align__(16) double sx[NL * ST_XSEG];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ double sh[ST_ROWS / 32];
  if (done && done[S_DONE] != 0.0) return;
  const int t = threadIdx.x;
  const int xpar = (int)(((uintptr_t)x >> 3) & 1);                              // bulk copies need 16-byte aligned sources
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // chunk -> (child, row range); chunks never straddle two child blocks
  struct Chunk { int ci; int64_t r0, r1, span0, span1; };
  int zero;                                                                       // opaque 0: keeps the prefetched metadata in per-thread
  asm volatile("mov.u32 %0, 0;" : "=r"(zero));                                   // registers (no early uniform-register move that would wait for the load)
  auto locate = [&](int64_t chunk, Chunk& C) {
    C.ci = 0; C.r0 = C.r1 = C.span0 = C.span1 = 0;
    if (chunk >= sp.nchunks) return;
    int ci = 0;
    while (ci + 1 < sp.nchild && chunk >= sp.ch[ci + 1].chunk0) ++ci;
    const SpmvChild& ch = sp.ch[ci];
    C.ci = ci;
    const int64_t v = chunk - ch.chunk0;
    C.r0 = ch.offset + (v < ch.skip_lo ? v : v + ch.skip_len) * ST_ROWS;
    C.r1 = (C.r0 + ST_ROWS < ch.offset + ch.size) ? C.r0 + ST_ROWS : ch.offset + ch.size;
    C.span0 = rowptr[C.r0 + zero]; C.span1 = rowptr[C.r1 + zero];
  };
  // a chunk is staged (values + x segments) when its x segments lie inside the vector and its value span fits the buffer
  auto staged_of = [&](const Chunk& C) {
    const int64_t a0 = C.span0 & ~(int64_t)1;
    return spmv_chunk_staged(sp.ch[C.ci], DIM, C.r0, C.r1, n) && (((C.span1 - a0) + 1) & ~(int64_t)1) <= ST_VCAP;
  };
  auto issue = [&](const Chunk& C) {                                              // thread 0 only
    const bool xst = spmv_chunk_staged(sp.ch[C.ci], DIM, C.r0, C.r1, n);
    const bool vst = staged_of(C);
    if (!xst) return;
    const int64_t a0 = C.span0 & ~(int64_t)1;
    const uint32_t vbytes = vst ? (uint32_t)((((C.span1 - a0) + 1) & ~(int64_t)1) * 8) : 0u;
    const int len = (int)(C.r1 - C.r0) + 2;
    uint32_t total = vbytes;
    if (xst) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int dy = l % 3 - 1, dz = (DIM == 3) ? l / 3 - 1 : 0;
        const int64_t g0 = C.r0 - 1 + dy * sp.ch[C.ci].bx + (int64_t)dz * sp.ch[C.ci].bxy;
        const int64_t al = g0 - ((g0 + xpar) & 1);
        total += (uint32_t)((((g0 + len) - al + 1) & ~(int64_t)1) * 8);
      }
    }
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(total) : "memory");
    if (vst)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sv)), "l"(val + a0),
                   "r"(vbytes), "r"(smem_u32(&bar))
                   : "memory");
    if (xst) {
#pragma unroll
      for (int l = 0; l < NL; ++l) {
        const int dy = l % 3 - 1, dz = (DIM == 3) ? l / 3 - 1 : 0;
        const int64_t g0 = C.r0 - 1 + dy * sp.ch[C.ci].bx + (int64_t)dz * sp.ch[C.ci].bxy;
        const int64_t al = g0 - ((g0 + xpar) & 1);
        const uint32_t bytes = (uint32_t)((((g0 + len) - al + 1) & ~(int64_t)1) * 8);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sx + l * ST_XSEG)),
                     "l"(x + al), "r"(bytes), "r"(smem_u32(&bar))
                     : "memory");
      }
    }
  };
  auto load_meta = [&](const Chunk& C, int64_t& s_, int64_t& e_, uint32_t& info_) {
    s_ = 0; e_ = 0; info_ = 0;
    if (C.r0 + t < C.r1) { s_ = rowptr[C.r0 + t]; e_ = rowptr[C.r0 + t + 1]; info_ = rowinfo[C.r0 + t]; }
  };
  uint32_t phase = 0;
  double d0 = 0.0, d1 = 0.0;
  Chunk C, Cn;
  int64_t s, e;
  uint32_t info;
  locate(blockIdx.x, C);
  load_meta(C, s, e, info);
  if (t == 0 && (int64_t)blockIdx.x < sp.nchunks) issue(C);
  for (int64_t chunk = blockIdx.x; chunk < sp.nchunks; chunk += gridDim.x) {     // block-uniform trip count
    // the block's next chunk: metadata loads are in flight while this one is summed
    int64_t s_n, e_n;
    uint32_t info_n;
    locate(chunk + gridDim.x, Cn);
    load_meta(Cn, s_n, e_n, info_n);
    const int64_t r = C.r0 + t;
    const bool live = r < C.r1;
    const int64_t a0 = C.span0 & ~(int64_t)1;
    const bool xst = spmv_chunk_staged(sp.ch[C.ci], DIM, C.r0, C.r1, n), vst = staged_of(C);      // block-uniform
    const int bx = sp.ch[C.ci].bx, bxy = sp.ch[C.ci].bxy;
    // affine rows of a staged chunk are summed here, every other row by k_spmv_rows (same predicate: k_spmv_rest_flags)
    const uint32_t m27 = (live && xst) ? spmv_affine_mask(DIM, info, e - s) : 0u;
    const bool affine = m27 != 0;
    const bool full = m27 == (1u << NS) - 1u;
    double sum = 0.0;
    if (xst) {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    if (affine) {
      const double* v = vst ? sv + (s - a0) : val + s;
      // segment l holds x[al_l ...], al_l = g0_l rounded down to a 16-byte boundary (g0_l = r0 - 1 + off_l): entry (t, dx) sits
      // at (g0_l - al_l) + t + dx + 1
      if (full) {
#pragma unroll
        for (int o = 0; o < NS; ++o) {
          const int dx = o % 3 - 1, l = o / 3, dy = l % 3 - 1, dz = (DIM == 3) ? l / 3 - 1 : 0;
          const int sh_l = (int)((C.r0 - 1 + dy * bx + (int64_t)dz * bxy + xpar) & 1);
          sum += v[o] * sx[l * ST_XSEG + sh_l + t + dx + 1];
        }
      } else {
#pragma unroll
        for (int o = 0; o < NS; ++o) {
          const int dx = o % 3 - 1, l = o / 3, dy = l % 3 - 1, dz = (DIM == 3) ? l / 3 - 1 : 0;
          const int sh_l = (int)((C.r0 - 1 + dy * bx + (int64_t)dz * bxy + xpar) & 1);
          if ((m27 >> o) & 1u) sum += v[__popc(m27 & ((1u << o) - 1u))] * sx[l * ST_XSEG + sh_l + t + dx + 1];
        }
      }
    }
    __syncthreads();                                                             // the staging buffers are free again:
    if (t == 0 && chunk + gridDim.x < sp.nchunks) issue(Cn);                      // next chunk's copies fly during the tail below
    if (affine) {
      y[r] = sum;
      if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
    }
    C = Cn; s = s_n; e = e_n; info = info_n;
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[pofs + blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + pofs + blockIdx.x] = s1;
    }
  }
}

// ---- shift-run product: structured numberings beyond Q1 (Qk entity groups) ----------------------------------------------------------
// On a structured mesh consecutive rows of an entity group mostly hold the SAME column list shifted by one (row r + 1 couples to
// column c + 1 wherever row r couples to c): along such a run the column indices need not be read at all.  The analysis is generic
// (it looks at the CSR pattern only, once per matrix): same[r] = row r has the length of row r - 1 and col[r][k] = col[r-1][k] + 1 for
// all k.  Maximal runs of at least SR_MINRUN rows are cut into chunks of <= 32 rows; a warp stages a chunk's values (one contiguous
// span, coalesced loads) in shared memory, lane t then sums row t in column order — entry k of the span at t m + k (row lengths of Qk
// rows are odd: conflict-free), x at col0[k] + t (coalesced) — so the product moves 8 bytes per stored entry instead of 12 and every
// lane carries a row (the vector-CSR kernel leaves half its lanes idle on 9 / 15 / 25-entry rows).  Rows outside the runs go through
// k_spmv_rows on a list.  Summation order = CSR order (the sequential order of the CPU product).
#define SR_MAXLEN 32
#define SR_MINRUN 8
__global__ void k_sr_same(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t* head)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  bool same = r > 0;
  if (same) {
    const int64_t s = rowptr[r], e = rowptr[r + 1], s0 = rowptr[r - 1];
    same = (e - s) == (s - s0) && (e - s) >= 1 && (e - s) <= SR_MAXLEN;
    for (int64_t k = 0; same && k < e - s; ++k) same = colidx[s + k] == colidx[s0 + k] + 1;
  }
  head[r] = same ? 0 : 1;
}
// run id of every row = (inclusive scan of head) - 1 = exclusive scan + head - 1; start[run] by the heads
__global__ void k_sr_starts(int64_t n, const int32_t* __restrict__ head, const int32_t* __restrict__ scan, int32_t* start)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n && head[r]) start[scan[r]] = (int32_t)r;
}
__global__ void k_sr_counts(int64_t nruns, int64_t n, const int32_t* __restrict__ start, const int64_t* __restrict__ rowptr, int32_t* nchunk, int* maxm)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nruns) return;
  const int64_t a = start[i], b = (i + 1 < nruns) ? start[i + 1] : n;
  const int64_t len = rowptr[a + 1] - rowptr[a];
  nchunk[i] = (b - a >= SR_MINRUN && len >= 1 && len <= SR_MAXLEN) ? (int32_t)((b - a + 31) / 32) : 0;
  if (nchunk[i] > 0) atomicMax(maxm, (int)len);
}
__global__ void k_sr_fill(int64_t n, const int32_t* __restrict__ head, const int32_t* __restrict__ scan, int64_t nruns, const int32_t* __restrict__ start,
                          const int32_t* __restrict__ nchunk, const int32_t* __restrict__ choff, const int64_t* __restrict__ rowptr, int4* chunks, int32_t* rest)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t run = scan[r] + head[r] - 1;
  const int64_t a = start[run], b = (run + 1 < nruns) ? start[run + 1] : n;
  const bool kept = nchunk[run] > 0;
  rest[r] = kept ? 0 : 1;
  if (kept && ((r - a) & 31) == 0) {                       // a chunk record: first row, rows | row length << 8, first entry (64 bits)
    const long long p0 = rowptr[r]; const int m = (int)(rowptr[r + 1] - p0);
    chunks[choff[run] + ((r - a) >> 5)] = make_int4((int)r, (int)((b - r) < 32 ? (b - r) : 32) | (m << 8), (int)(p0 & 0xffffffffll), (int)(p0 >> 32));
  }
}
#define SR_WARPS 4
#define SR_STAGES 4
// Values travel through a ring of SR_STAGES bulk-async copies per warp (ncu of the first, double-buffered version: 12 warps per SM with
// one 6 KB copy in flight each = 77 KB per SM in flight, 7 300 cycles per chunk, DRAM 32 %: under load a copy takes microseconds, so the
// bytes in flight bound the rate — the Q1 stencil kernel keeps ~150 KB per SM in flight).  Slots are sized by the longest row of the
// matrix' runs (`slot` doubles, even), so 2 blocks x 4 warps x 4 stages fit an SM for 25-entry rows.
template <int DOTS>
__global__ void __launch_bounds__(32 * SR_WARPS) k_spmv_runs(int64_t nchunks, const int4* __restrict__ chunks, int slot, const int32_t* __restrict__ colidx,
                                                            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                            const double* __restrict__ a, const uint8_t* __restrict__ owned, double* partial, int pofs,
                                                            const double* __restrict__ done)
{
  extern __shared__ __align__(16) double sr_smem[];                  // [SR_WARPS][SR_STAGES][slot] values, then [SR_WARPS][SR_STAGES] mbarriers
  __shared__ double sh[8];
  if (done && done[S_DONE] != 0.0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* buf = sr_smem + (size_t)warp * SR_STAGES * slot;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sr_smem + (size_t)SR_WARPS * SR_STAGES * slot) + warp * SR_STAGES;
  if (lane == 0) {
    for (int i = 0; i < SR_STAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  struct Ck { int64_t r0, p0; int nr, m; };
  const int4 none = make_int4(0, 0, 0, 0);
  auto decode = [](const int4 q, Ck& C) {
    C.r0 = q.x; C.nr = q.y & 255; C.m = q.y >> 8; C.p0 = (int64_t)(((unsigned long long)(unsigned)q.w << 32) | (unsigned)q.z);
  };
  auto issue = [&](const int4 q, int st) {                             // lane 0 only
    Ck C; decode(q, C);
    if (C.nr == 0) return;
    const int64_t a0 = C.p0 & ~(int64_t)1;
    const uint32_t bytes = (uint32_t)((((C.p0 + (int64_t)C.nr * C.m) - a0 + 1) & ~(int64_t)1) * 8);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + st)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + (size_t)st * slot)),
                 "l"(val + a0), "r"(bytes), "r"(smem_u32(bar + st))
                 : "memory");
  };
  double d0 = 0.0, d1 = 0.0;
  const int64_t step = (int64_t)gridDim.x * SR_WARPS;
  const int64_t ck0 = blockIdx.x * (int64_t)SR_WARPS + warp;
  // prologue: the first SR_STAGES - 1 chunks of this warp are in flight; chunk records are read SR_STAGES iterations ahead
  for (int i = 0; i < SR_STAGES - 1; ++i) {
    const int64_t cki = ck0 + i * step;
    if (lane == 0 && cki < nchunks) issue(chunks[cki], i);
  }
  Ck C;
  decode(ck0 < nchunks ? chunks[ck0] : none, C);
  int4 qnext = ck0 + step < nchunks ? chunks[ck0 + step] : none;       // record of the chunk summed next
  int4 qissue = ck0 + (SR_STAGES - 1) * step < nchunks ? chunks[ck0 + (SR_STAGES - 1) * step] : none;   // record of the chunk issued next
  int cmine = (C.nr > 0 && lane < C.m) ? __ldg(colidx + C.p0 + lane) : 0;   // the run's column list: first row of the chunk, shifted by the lane
  int st = 0;
  uint32_t phases = 0u;                                                 // bit s = parity of stage s
  for (int64_t ck = ck0; ck < nchunks; ck += step) {
    // stage (st + SR_STAGES - 1) % SR_STAGES was consumed in the previous iteration (__syncwarp below): refill it
    if (lane == 0) issue(qissue, (st + SR_STAGES - 1) % SR_STAGES);
    qissue = ck + SR_STAGES * step < nchunks ? chunks[ck + SR_STAGES * step] : none;
    Ck Cn; decode(qnext, Cn);
    const int cnext = (Cn.nr > 0 && lane < Cn.m) ? __ldg(colidx + Cn.p0 + lane) : 0;
    qnext = ck + 2 * step < nchunks ? chunks[ck + 2 * step] : none;
    const bool live = lane < C.nr;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar + st)), "r"((phases >> st) & 1u) : "memory");
    phases ^= 1u << st;
    const double* mine = buf + (size_t)st * slot + (C.p0 & 1) + lane * C.m;
    double sum = 0.0;
    int k = 0;
    for (; k + 8 <= C.m; k += 8) {                                      // eight x loads in flight, summed in column order
      double xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int cj = __shfl_sync(0xffffffffu, cmine, k + j); xv[j] = live ? x[(int64_t)cj + lane] : 0.0; }
      if (live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += mine[k + j] * xv[j];
      }
    }
    for (; k < C.m; ++k) {
      const int c0 = __shfl_sync(0xffffffffu, cmine, k);
      if (live) sum += mine[k] * x[(int64_t)c0 + lane];
    }
    if (live) {
      const int64_t r = C.r0 + lane;
      y[r] = sum;
      if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
    }
    __syncwarp();
    C = Cn; cmine = cnext;
    st = (st + 1) % SR_STAGES;
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[pofs + blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + pofs + blockIdx.x] = s1;
    }
  }
}
// The same chunks without shared memory: 8 lanes per row as in k_spmv_vec (full occupancy: the product is bound by load latency, not by
// bytes), but the column list of the chunk's first row lives in the lanes (one coalesced load per 32 rows) — no column-index traffic and
// no colidx -> x dependency: the value and x loads of 8 rows x 4 entries per lane are independent and issued together.
template <int DOTS>
__global__ void __launch_bounds__(256) k_spmv_runs_vec(int64_t nchunks, const int4* __restrict__ chunks, const int32_t* __restrict__ colidx,
                                                       const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y,
                                                       const double* __restrict__ a, const uint8_t* __restrict__ owned, double* partial, int pofs,
                                                       const double* __restrict__ done)
{
  __shared__ double sh[8];
  if (done && done[S_DONE] != 0.0) return;
  const int lane = threadIdx.x & 31, sub = lane & 7, g = lane >> 3;
  double d0 = 0.0, d1 = 0.0;
  const int64_t step = (int64_t)gridDim.x * 8;
  for (int64_t ck = blockIdx.x * 8ll + (threadIdx.x >> 5); ck < nchunks; ck += step) {
    const int4 q = chunks[ck];
    const int64_t r0 = q.x; const int nr = q.y & 255, m = q.y >> 8;
    const int64_t p0 = (int64_t)(((unsigned long long)(unsigned)q.w << 32) | (unsigned)q.z);
    const int cmine = (lane < m) ? __ldg(colidx + p0 + lane) : 0;
    int c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = __shfl_sync(0xffffffffu, cmine, (sub + 8 * j) & 31);
#pragma unroll 2
    for (int t = 0; t < 32; t += 4) {                                 // 4 rows per step, 8 lanes each
      if (t >= nr) break;                                               // warp-uniform
      const int row = t + g;
      const bool live = row < nr;
      const double* v = val + p0 + (int64_t)row * m;
      double vv[4], xx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool on = live && sub + 8 * j < m;
        vv[j] = on ? ld_stream_f64(v + sub + 8 * j) : 0.0;
        xx[j] = on ? x[(int64_t)c[j] + row] : 0.0;
      }
      double sum = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += vv[j] * xx[j];
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, 8);
      if (live && sub == 0) {
        const int64_t r = r0 + row;
        y[r] = sum;
        if (DOTS >= 1 && (!owned || owned[r])) { d0 += sum * a[r]; if (DOTS >= 2) d1 += sum * sum; }
      }
    }
  }
  if (DOTS >= 1) {
    double s0 = block_sum(d0, sh);
    if (threadIdx.x == 0) partial[pofs + blockIdx.x] = s0;
    if (DOTS >= 2) {
      double s1 = block_sum(d1, sh);
      if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + pofs + blockIdx.x] = s1;
    }
  }
}
static int64_t sr_scan(Ctx* c, const int32_t* d_in, int32_t* d_out, int64_t n)
{
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_in, d_out, (int)n, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_in, d_out, (int)n, c->stream)); c->launches++;
  int32_t ls = 0, lf = 0;
  MDB_CUDA(cudaMemcpyAsync(&ls, d_out + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&lf, d_in + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_t);
  return (int64_t)ls + lf;
}
// first product with this matrix: runs, chunks and the list of the remaining rows.  Leaves n_sr_chunks = 0 when fewer than 70 % of the
// rows sit in runs (the vector-CSR kernel then takes the whole matrix).
static void sr_analyse(Ctx* c, Matrix& M)
{
  const int64_t n = c->ndof;
  M.n_sr_chunks = 0;
  if (n < 2 || n >= (int64_t)2147483647 || M.max_row < 1) return;
  int32_t* d_head = dalloc<int32_t>(n); int32_t* d_scan = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_sr_same, cdiv(n, 256), 256, 0, n, M.d_rowptr, M.d_colidx, d_head);
  const int64_t nruns = sr_scan(c, d_head, d_scan, n);
  int32_t* d_start = dalloc<int32_t>(nruns); int32_t* d_nchunk = dalloc<int32_t>(nruns); int32_t* d_choff = dalloc<int32_t>(nruns);
  MDB_LAUNCH(c, k_sr_starts, cdiv(n, 256), 256, 0, n, d_head, d_scan, d_start);
  int* d_maxm = dalloc<int>(1);
  MDB_CUDA(cudaMemsetAsync(d_maxm, 0, sizeof(int), c->stream));
  MDB_LAUNCH(c, k_sr_counts, cdiv(nruns, 256), 256, 0, nruns, n, d_start, M.d_rowptr, d_nchunk, d_maxm);
  int maxm = 0;
  MDB_CUDA(cudaMemcpyAsync(&maxm, d_maxm, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  const int64_t nchunks = sr_scan(c, d_nchunk, d_choff, nruns);
  int4* d_chunks = dalloc<int4>(nchunks > 0 ? nchunks : 1);
  int32_t* d_rest = dalloc<int32_t>(n); int32_t* d_rscan = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_sr_fill, cdiv(n, 256), 256, 0, n, d_head, d_scan, nruns, d_start, d_nchunk, d_choff, M.d_rowptr, d_chunks, d_rest);
  const int64_t nrest = sr_scan(c, d_rest, d_rscan, n);
  if (nchunks > 0 && (n - nrest) * 10 >= n * 7) {
    dfree(M.d_spmv_rest);
    M.d_spmv_rest = dalloc<int32_t>(nrest > 0 ? nrest : 1);
    if (nrest > 0) MDB_LAUNCH(c, k_spmv_rest_compact, cdiv(n, 256), 256, 0, n, d_rest, d_rscan, M.d_spmv_rest);
    M.n_spmv_rest = nrest;
    M.d_sr_chunks = d_chunks; M.n_sr_chunks = nchunks; M.sr_maxm = maxm;
    d_chunks = nullptr;
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  if (getenv("MDB_SR_DEBUG")) fprintf(stderr, "shift-run analysis: %lld rows, %lld runs, %lld chunks, %lld rows outside -> %s\n", (long long)n, (long long)nruns,
                                      (long long)nchunks, (long long)nrest, M.n_sr_chunks > 0 ? "run kernel" : "vector-CSR kernel");
  dfree(d_head); dfree(d_scan); dfree(d_start); dfree(d_nchunk); dfree(d_choff); dfree(d_rest); dfree(d_rscan); dfree(d_maxm); if (d_chunks) dfree(d_chunks);
}

static int spmv_variant()
{
  static int v = -1;
  if (v < 0) { const char* e = getenv("MDB_SPMV_VARIANT"); v = e ? atoi(e) : 4; }     // 4 = stencil-aware (Q1 spaces only)
  return v;
}

// part: 0 = the whole product; on a partition 1 = the rows that read no ghost entry (issued while the halo is in flight), 2 = the
// rest (after the halo wait).  *nblocks is the number of partial sums written so far: part 2 appends to part 1's.
static bool spmv_can_split(const Ctx* c)
{
  if (spmv_variant() != 4 || c->um || c->part_axis < 0) return false;
  for (const Child& ch : c->children) if (ch.k != 1 || ch.dg) return false;
  return true;
}

template <int DOTS>
static void launch_spmv(Ctx* c, const Matrix& M, const double* x, double* y, const double* a, double* partial, int* nblocks, const double* done, int part = 0)
{
  const int64_t n = c->ndof;
  const size_t smem = (size_t)SPMV_ROWS * M.max_row * sizeof(double);
  int variant = spmv_variant();
  if (smem > 100 * 1024 && variant == 0) variant = 1;      // very long rows: no staging
  if (variant == 0) {
    static size_t configured[3] = {0, 0, 0};
    if (smem > 48 * 1024 && smem > configured[DOTS]) {
      MDB_CUDA(cudaFuncSetAttribute(k_spmv<DOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured[DOTS] = smem;
    }
    const int nchunks = cdiv(n, SPMV_ROWS);
    int per_sm = (int)(200 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / SPMV_THREADS) per_sm = 2048 / SPMV_THREADS;
    int grid = 148 * per_sm;
    if (grid > MAX_PARTIAL_BLOCKS) grid = MAX_PARTIAL_BLOCKS;
    if (grid > nchunks) grid = nchunks;
    if (nblocks) *nblocks = grid;
    MDB_LAUNCH(c, k_spmv<DOTS>, grid, SPMV_THREADS, smem, n, nchunks, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, done);
    return;
  }
  if (variant == 4) {
    bool q1 = c->um == nullptr;               // the stencil kernel needs lattice blocks
    for (const Child& ch : c->children) q1 = q1 && ch.k == 1 && !ch.dg;
    if (q1) {
      SpmvSpace sp; sp.nchild = (int)c->children.size(); sp.nchunks = 0;
      const int dim = c->mesh.dim;
      for (int i = 0; i < sp.nchild; ++i) {
        const Child& ch = c->children[i];
        int64_t vol = 1;
        for (int d = 0; d < dim; ++d) vol *= ch.lat_hi[d] - ch.lat_lo[d] + 1;
        sp.ch[i].offset = ch.offset; sp.ch[i].size = ch.size; sp.ch[i].box = (ch.size > 0 && vol == ch.size) ? 1 : 0;
        sp.ch[i].bx = ch.lat_hi[0] - ch.lat_lo[0] + 1; sp.ch[i].bxy = sp.ch[i].bx * (ch.lat_hi[1] - ch.lat_lo[1] + 1);
        const int64_t all = (ch.size + ST_ROWS - 1) / ST_ROWS;
        // interior chunks: rows of the lattice planes [zlo, zhi) along the cut axis, whose stencils stay clear of the ghost planes
        int64_t ia = 0, ib = 0;
        if (part != 0 && sp.ch[i].box && c->part_axis == dim - 1) {
          const int64_t nz = ch.lat_hi[dim - 1] - ch.lat_lo[dim - 1] + 1, plane = (dim == 3) ? sp.ch[i].bxy : sp.ch[i].bx;
          const int64_t zlo = c->peer_lower ? 2 : 0, zhi = c->peer_upper ? nz - 2 : nz;
          if (zhi > zlo) { ia = (zlo * plane + ST_ROWS - 1) / ST_ROWS; ib = (zhi * plane) / ST_ROWS; if (ib < ia) ib = ia; if (ib > all) ib = all; }
        }
        sp.ch[i].chunk0 = sp.nchunks;
        if (part == 1) { sp.ch[i].skip_lo = 0; sp.ch[i].skip_len = ia; sp.nchunks += ib - ia; }
        else if (part == 2) { sp.ch[i].skip_lo = ia; sp.ch[i].skip_len = ib - ia; sp.nchunks += all - (ib - ia); }
        else { sp.ch[i].skip_lo = 0; sp.ch[i].skip_len = 0; sp.nchunks += all; }
      }
      if (M.n_spmv_rest < 0) {                // first product with this matrix: the rows the stencil kernel leaves out
        Matrix& Mm = const_cast<Matrix&>(M);
        int32_t* d_flag = dalloc<int32_t>(n);
        int32_t* d_scan = dalloc<int32_t>(n);
        MDB_LAUNCH(c, k_spmv_rest_flags, cdiv(n, 256), 256, 0, n, sp, dim, M.d_rowptr, M.d_rowinfo, d_flag);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_scan, (int)n, c->stream);
        void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
        MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_flag, d_scan, (int)n, c->stream)); c->launches++;
        int32_t ls = 0, lf = 0;
        MDB_CUDA(cudaMemcpyAsync(&ls, d_scan + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        MDB_CUDA(cudaMemcpyAsync(&lf, d_flag + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        MDB_CUDA(cudaStreamSynchronize(c->stream));
        Mm.d_spmv_rest = dalloc<int32_t>((int64_t)ls + lf);
        MDB_LAUNCH(c, k_spmv_rest_compact, cdiv(n, 256), 256, 0, n, d_flag, d_scan, Mm.d_spmv_rest);
        MDB_CUDA(cudaStreamSynchronize(c->stream));
        Mm.n_spmv_rest = (int64_t)ls + lf;
        cudaFree(d_t); dfree(d_flag); dfree(d_scan);
      }
      const int pofs = (part == 2 && nblocks) ? *nblocks : 0;
      int grid = 148 * ST_BLOCKS;
      if (part == 2) grid = 148 * 2;                                   // the boundary planes are a few hundred chunks
      if (grid > sp.nchunks) grid = (int)sp.nchunks;
      int grid2 = (part == 1) ? 0 : 148 * 8;                           // the row list (interface rows, vector ends) may read ghosts: with the rest
      if (grid2 > cdiv(M.n_spmv_rest, 32)) grid2 = cdiv(M.n_spmv_rest, 32);
      if (grid2 > MAX_PARTIAL_BLOCKS - pofs - grid) grid2 = MAX_PARTIAL_BLOCKS - pofs - grid;      // one partial sum per block (grid-stride kernels)
      if (nblocks) *nblocks = pofs + grid + grid2;
      MDB_REQUIRE(pofs + grid + grid2 <= MAX_PARTIAL_BLOCKS, MDB_EINVAL, "spmv: too many partial sums");
      if (grid > 0) {
        if (dim == 3) MDB_LAUNCH(c, (k_spmv_stencil<3, DOTS>), grid, ST_ROWS, 0, n, sp, M.d_rowptr, M.d_rowinfo, M.d_val, x, y, a, c->d_owned, partial, pofs, done);
        else MDB_LAUNCH(c, (k_spmv_stencil<2, DOTS>), grid, ST_ROWS, 0, n, sp, M.d_rowptr, M.d_rowinfo, M.d_val, x, y, a, c->d_owned, partial, pofs, done);
      }
      if (grid2 > 0)
        MDB_LAUNCH(c, k_spmv_rows<DOTS>, grid2, 256, 0, M.n_spmv_rest, M.d_spmv_rest, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial,
                   pofs + grid, done);
      return;
    }
    // structured numberings beyond Q1 (Qk entity groups): the shift-run kernel where the pattern has the runs for it
    // MEASURED (B200, Q2 | Q2 on 2048^2 cells, 16.8 M rows, 2.7e8 entries, profiles/r02/qk_spmv_runs_r4.md): 99.9 % of the rows sit in
    // runs; the product alone goes 0.99 -> 0.78 ms with k_spmv_runs_vec (x warm in L2 between repetitions), but inside CG, where four
    // other vectors pass through L2 between two products, the iteration stays at 1.28 -> 1.31 ms: the product is bound by the latency
    // of the x gathers, not by the column-index bytes.  The shared-memory ring kernel loses outright (1.12 - 2.58 ms: 8 - 12 warps per SM
    // cannot hide the x loads; ncu: every stall a long scoreboard).  Both stay as switches; the default is the vector-CSR kernel.
    static const char* runs_env = getenv("MDB_SPMV_RUNS");                  // unset: off; "staged": ring kernel; anything else: k_spmv_runs_vec
    const bool no_runs = runs_env == nullptr;
    bool qk = c->um == nullptr && !no_runs && M.max_row <= 4 * SR_MAXLEN;
    for (const Child& ch : c->children) qk = qk && !ch.dg;
    if (qk) {
      Matrix& Mm = const_cast<Matrix&>(M);
      if (M.n_sr_chunks < 0) sr_analyse(c, Mm);
      if (M.n_sr_chunks > 0) {
        const bool staged = std::strcmp(runs_env, "staged") == 0;            // A/B switch: the shared-memory ring kernel
        int grid = 148 * 8;
        size_t smem = 0; int slot = 0;
        if (staged) {
          slot = (32 * M.sr_maxm + 2 + 1) & ~1;                          // doubles per stage: a chunk's span from the 16-byte boundary below
          smem = (size_t)SR_WARPS * SR_STAGES * slot * sizeof(double) + SR_WARPS * SR_STAGES * sizeof(unsigned long long);
          static size_t configured[3] = {0, 0, 0};
          if (smem > configured[DOTS]) {
            MDB_CUDA(cudaFuncSetAttribute(k_spmv_runs<DOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[DOTS] = smem;
          }
          int per_sm = (int)((226 * 1024) / (smem + 1024 + 128));
          if (per_sm < 1) per_sm = 1;
          grid = 148 * per_sm;                                           // resident blocks only: every warp keeps its ring full
          if (grid > cdiv(M.n_sr_chunks, SR_WARPS)) grid = cdiv(M.n_sr_chunks, SR_WARPS);
        } else if (grid > cdiv(M.n_sr_chunks, 8)) grid = cdiv(M.n_sr_chunks, 8);
        int grid2 = 148 * 6;
        if (grid2 > cdiv(M.n_spmv_rest, 32)) grid2 = cdiv(M.n_spmv_rest, 32);
        if (grid2 > MAX_PARTIAL_BLOCKS - grid) grid2 = MAX_PARTIAL_BLOCKS - grid;      // one partial sum per block (grid-stride kernels)
        if (nblocks) *nblocks = grid + grid2;
        if (staged)
          MDB_LAUNCH(c, k_spmv_runs<DOTS>, grid, 32 * SR_WARPS, smem, M.n_sr_chunks, M.d_sr_chunks, slot, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, 0, done);
        else
          MDB_LAUNCH(c, k_spmv_runs_vec<DOTS>, grid, 256, 0, M.n_sr_chunks, M.d_sr_chunks, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, 0, done);
        if (grid2 > 0)
          MDB_LAUNCH(c, k_spmv_rows<DOTS>, grid2, 256, 0, M.n_spmv_rest, M.d_spmv_rest, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, grid, done);
        return;
      }
    }
    variant = 1;
  }
  const int lpr = (variant == 2) ? 4 : (variant == 3) ? 16 : 8;
  int grid = 148 * 8;
  const int64_t need = cdiv(n, 256 / lpr);
  if (grid > need) grid = (int)need;
  if (nblocks) *nblocks = grid;
  if (lpr == 4) MDB_LAUNCH(c, (k_spmv_vec<DOTS, 4>), grid, 256, 0, n, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, done);
  else if (lpr == 16) MDB_LAUNCH(c, (k_spmv_vec<DOTS, 16>), grid, 256, 0, n, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, done);
  else MDB_LAUNCH(c, (k_spmv_vec<DOTS, 8>), grid, 256, 0, n, M.d_rowptr, M.d_colidx, M.d_val, x, y, a, c->d_owned, partial, done);
}

// y = A x on a partition: the owners' boundary planes travel while the interior rows are summed
template <int DOTS>
static void spmv_halo(Ctx* c, const Matrix& M, double* x, double* y, const double* a, double* partial, int* nblocks, double* s)
{
  // Measured on 2 x B200 (profiles/r02/README.md): every rank pushes right after an all-reduce, so the planes arrive within microseconds
  // and there is no waiting time to hide; splitting the product costs a second, poorly filled launch (0.694 against 0.682 ms per
  // product, CG 984 against 994 iterations/s).  The split is therefore opt-in (MDB_HALO_OVERLAP=1: for fabrics with slower neighbours).
  static const bool overlap = getenv("MDB_HALO_OVERLAP") != nullptr;
  if (!overlap || !spmv_can_split(c)) {
    halo_exchange(c, x, s ? s + S_DONE : nullptr);
    launch_spmv<DOTS>(c, M, x, y, a, partial, nblocks, s);
    return;
  }
  halo_push(c, x);
  launch_spmv<DOTS>(c, M, x, y, a, partial, nblocks, s, 1);
  halo_wait(c, x, s ? s + S_DONE : nullptr);
  launch_spmv<DOTS>(c, M, x, y, a, partial, nblocks, s, 2);
}

void spmv(Ctx* c, const Matrix& M, const double* d_x, double* d_y)
{
  launch_spmv<0>(c, M, d_x, d_y, nullptr, nullptr, nullptr, nullptr);
}
void spmv_partitioned(Ctx* c, const Matrix& M, double* d_x, double* d_y)
{
  int nb = 0;
  spmv_halo<0>(c, M, d_x, d_y, nullptr, nullptr, &nb, nullptr);
}

// ---- scalar helpers (single block) -------------------------------------------------------------------
// out[slot_k] = sum of partial[k][0..nblocks) for k < nvals, in a fixed order (deterministic)
__global__ void k_reduce_partials(const double* __restrict__ partial, int nblocks, int nvals, double* scal, int slot0, int slot1, int slot2)
{
  __shared__ double sh[VEC_THREADS / 32];
  for (int k = 0; k < nvals; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partial[k * MAX_PARTIAL_BLOCKS + i];
    double s = block_sum(v, sh);
    if (threadIdx.x == 0) scal[k == 0 ? slot0 : (k == 1 ? slot1 : slot2)] = s;
    __syncthreads();
  }
}
// the same reduction ending in the peer-memory all-reduce over the ranks (sr_exchange, common.cuh): one launch instead of a reduction
// kernel + an NCCL all-reduce
__global__ void k_reduce_partials_ranks(const double* __restrict__ partial, int nblocks, int nvals, double* scal, int slot0, int slot1, int slot2, SrSpec R)
{
  __shared__ double sh[VEC_THREADS / 32];
  __shared__ double loc[SR_MAXV];
  for (int k = 0; k < nvals; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) v += partial[k * MAX_PARTIAL_BLOCKS + i];
    double s = block_sum(v, sh);
    if (threadIdx.x == 0) loc[k] = s;
    __syncthreads();
  }
  sr_exchange(R, loc, nvals);
  if (threadIdx.x == 0)
    for (int k = 0; k < nvals; ++k) scal[k == 0 ? slot0 : (k == 1 ? slot1 : slot2)] = loc[k];
}

__global__ void k_extract_dinv(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ diag, const double* __restrict__ val,
                               double* dinv, int jacobi)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d = 1.0;
  if (jacobi && diag[i] >= 0) { double a = val[rowptr[i] + diag[i]]; d = (a != 0.0) ? 1.0 / a : 1.0; }
  dinv[i] = d;
}

// ---- CG ([UPSTREAM] dune-istl CGSolver::apply) -------------------------------------------------------
// init: x = 0, r = b, p = 0; partial0 = sum dinv r^2 (rho), partial1 = sum r^2
__global__ void k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ dinv, const uint8_t* __restrict__ owned,
                          double* x, double* r, double* p, double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  double a0 = 0.0, a1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double ri = b[i];
    x[i] = 0.0; r[i] = ri; p[i] = 0.0;
    if (!owned || owned[i]) { a0 += dinv[i] * ri * ri; a1 += ri * ri; }
  }
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
  double s1 = block_sum(a1, sh); if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + blockIdx.x] = s1;
}

__global__ void k_cg_init_scalars(double* s, double reduction)
{
  // S_RHONEW and S_RR were just reduced
  s[S_RHO] = s[S_RHONEW];
  s[S_RR0] = s[S_RR]; s[S_RHOLAST] = 1.0; s[S_ITER] = 0.0; s[S_RED2] = reduction * reduction; s[S_BREAK] = 0.0;
  s[S_BETA] = s[S_RHO] / 1.0;
  s[S_DONE] = (s[S_RR] == 0.0) ? 1.0 : 0.0;
}

// p = beta p + dinv r      (q = M^-1 r ; p *= beta ; p += q)
__global__ void k_cg_update_p(int64_t n, const double* __restrict__ s, const double* __restrict__ dinv, const double* __restrict__ r, double* p)
{
  if (s[S_DONE] != 0.0) return;
  const double beta = s[S_BETA];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = p[i] * beta + dinv[i] * r[i];
}

// lambda = rho / (p.q)
__global__ void k_cg_derive_a(double* s) { if (s[S_DONE] != 0.0) return; s[S_LAMBDA] = s[S_RHO] / s[S_PQ]; }

// x += lambda p ; r -= lambda q ; partial0 = sum dinv r^2 (next rho), partial1 = sum r^2
__global__ void k_cg_update_xr(int64_t n, const double* __restrict__ s, const double* __restrict__ dinv, const double* __restrict__ p,
                               const double* __restrict__ q, const uint8_t* __restrict__ owned, double* x, double* r, double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  if (s[S_DONE] != 0.0) return;
  const double lambda = s[S_LAMBDA];
  double a0 = 0.0, a1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += lambda * p[i];
    double ri = r[i] - lambda * q[i];
    r[i] = ri;
    if (!owned || owned[i]) { a0 += dinv[i] * ri * ri; a1 += ri * ri; }
  }
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
  double s1 = block_sum(a1, sh); if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + blockIdx.x] = s1;
}

// rholast = rho ; rho = new ; beta = rho/rholast ; iteration count ; convergence test  def <= reduction * def0
__global__ void k_cg_derive_b(double* s)
{
  if (s[S_DONE] != 0.0) return;
  s[S_ITER] += 1.0;
  s[S_RHOLAST] = s[S_RHO];
  s[S_RHO] = s[S_RHONEW];
  s[S_BETA] = s[S_RHO] / s[S_RHOLAST];
  if (s[S_RR] <= s[S_RED2] * s[S_RR0]) s[S_DONE] = 1.0;
}

// ---- BiCGSTAB ([UPSTREAM] dune-istl BiCGSTABSolver::apply, right-preconditioned as ISTL does) --------
__global__ void k_bi_init(int64_t n, const double* __restrict__ b, const uint8_t* __restrict__ owned, double* x, double* r, double* rt,
                          double* p, double* v, double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  double a0 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double ri = b[i];
    x[i] = 0.0; r[i] = ri; rt[i] = ri; p[i] = 0.0; v[i] = 0.0;
    if (!owned || owned[i]) a0 += ri * ri;
  }
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
}
__global__ void k_bi_init_scalars(double* s, double reduction)
{
  // S_RR reduced; rho_new = rt.r = r.r
  s[S_RR0] = s[S_RR]; s[S_RHONEW] = s[S_RR]; s[S_RHO] = 1.0; s[S_ALPHA] = 1.0; s[S_OMEGA] = 1.0; s[S_ITER] = 0.0;
  s[S_RED2] = reduction * reduction; s[S_HALF] = 0.0; s[S_BREAK] = 0.0;
  s[S_BETA] = 0.0;          // first iteration: p = r
  s[S_DONE] = (s[S_RR] == 0.0) ? 1.0 : 0.0;
}
// p = r + beta (p - omega v) ; y = dinv p
__global__ void k_bi_update_p(int64_t n, const double* __restrict__ s, const double* __restrict__ dinv, const double* __restrict__ r,
                              const double* __restrict__ v, double* p, double* y)
{
  if (s[S_DONE] != 0.0) return;
  const double beta = s[S_BETA], omega = s[S_OMEGA];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double pi = r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi; y[i] = dinv[i] * pi;
  }
}
// [UPSTREAM] BiCGSTABSolver::apply aborts with an ISTLError when |rho|, |h| or |omega| fall below EPSILON = 1e-80 (and a NaN fails its
// comparisons likewise): here the loop stops (S_DONE) with S_BREAK set and the solve reports "not converged" instead of iterating on NaNs
#define MDB_BICG_EPS 1e-80
__global__ void k_bi_derive_alpha(double* s)
{
  if (s[S_DONE] != 0.0) return;
  if (!(fabs(s[S_RHONEW]) > MDB_BICG_EPS) || !(fabs(s[S_H]) > MDB_BICG_EPS)) { s[S_DONE] = 1.0; s[S_BREAK] = 1.0; return; }
  s[S_ALPHA] = s[S_RHONEW] / s[S_H];
}
// x += alpha y ; r -= alpha v ; y2 = dinv r ; partial0 = sum r^2
__global__ void k_bi_update_s(int64_t n, const double* __restrict__ s, const double* __restrict__ dinv, const double* __restrict__ y,
                              const double* __restrict__ v, const uint8_t* __restrict__ owned, double* x, double* r, double* y2, double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  if (s[S_DONE] != 0.0) return;
  const double alpha = s[S_ALPHA];
  double a0 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * y[i];
    double ri = r[i] - alpha * v[i];
    r[i] = ri; y2[i] = dinv[i] * ri;
    if (!owned || owned[i]) a0 += ri * ri;
  }
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
}
// half-iteration convergence test (ISTL tests after the first update too)
__global__ void k_bi_derive_half(double* s)
{
  if (s[S_DONE] != 0.0) return;
  if (s[S_RR] <= s[S_RED2] * s[S_RR0]) { s[S_DONE] = 1.0; s[S_ITER] += 1.0; s[S_HALF] = 1.0; }
}
__global__ void k_bi_derive_omega(double* s)
{
  if (s[S_DONE] != 0.0) return;
  const double omega = s[S_TR] / s[S_TT];
  if (!(fabs(omega) > MDB_BICG_EPS)) { s[S_DONE] = 1.0; s[S_BREAK] = 1.0; return; }
  s[S_OMEGA] = omega;
}
// x += omega y2 ; r -= omega t ; partial0 = sum r^2 ; partial1 = sum rt r
__global__ void k_bi_update_x(int64_t n, const double* __restrict__ s, const double* __restrict__ y2, const double* __restrict__ t,
                              const double* __restrict__ rt, const uint8_t* __restrict__ owned, double* x, double* r, double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  if (s[S_DONE] != 0.0) return;
  const double omega = s[S_OMEGA];
  double a0 = 0.0, a1 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += omega * y2[i];
    double ri = r[i] - omega * t[i];
    r[i] = ri;
    if (!owned || owned[i]) { a0 += ri * ri; a1 += rt[i] * ri; }
  }
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
  double s1 = block_sum(a1, sh); if (threadIdx.x == 0) partial[MAX_PARTIAL_BLOCKS + blockIdx.x] = s1;
}
__global__ void k_bi_derive_end(double* s)
{
  if (s[S_DONE] != 0.0) return;
  s[S_ITER] += 1.0;
  // S_RHONEW currently holds rho of this iteration; S_TR slot reused: new rt.r was reduced into S_PQ
  double rho_old = s[S_RHONEW];
  s[S_RHO] = rho_old;
  s[S_RHONEW] = s[S_PQ];
  s[S_BETA] = (s[S_RHONEW] / s[S_RHO]) * (s[S_ALPHA] / s[S_OMEGA]);
  if (s[S_RR] <= s[S_RED2] * s[S_RR0]) s[S_DONE] = 1.0;
}

// partial0 = sum a_i b_i over owned rows
__global__ void k_dot(int64_t n, const double* __restrict__ s, const double* __restrict__ a, const double* __restrict__ b, const uint8_t* __restrict__ owned,
                      double* partial)
{
  __shared__ double sh[VEC_THREADS / 32];
  if (s[S_DONE] != 0.0) return;
  double a0 = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!owned || owned[i]) a0 += a[i] * b[i];
  double s0 = block_sum(a0, sh); if (threadIdx.x == 0) partial[blockIdx.x] = s0;
}

static void reduce(Ctx* c, int nblocks, int nvals, int s0, int s1 = 0, int s2 = 0)
{
  MDB_LAUNCH(c, k_reduce_partials, 1, VEC_THREADS, 0, c->d_partial, nblocks, nvals, c->d_scal, s0, s1, s2);
}
// block partials -> scalars s0 (, s1, s2), summed over the ranks of the communicator
static void reduce_ranks(Ctx* c, bool multi, int nblocks, int nvals, int s0, int s1 = 0, int s2 = 0)
{
  if (multi && sring_connected(c)) {
    MDB_LAUNCH(c, k_reduce_partials_ranks, 1, VEC_THREADS, 0, c->d_partial, nblocks, nvals, c->d_scal, s0, s1, s2, sring_next(c, c->d_scal + S_DONE));
    return;
  }
  reduce(c, nblocks, nvals, s0, s1, s2);
  if (multi) {
    // NCCL reduces a contiguous range
    MDB_REQUIRE(nvals == 1 || (s1 == s0 + 1 && (nvals == 2 || s2 == s0 + 2)), MDB_EINVAL, "reduce_ranks: scalar slots must be adjacent");
    allreduce_scalars(c, c->d_scal + s0, nvals);
  }
}

static void ensure_work(Ctx* c, int nvecs)
{
  if (c->work_n >= c->ndof && c->work_vecs >= nvecs) return;
  dfree(c->d_work);
  c->d_work = dalloc<double>((size_t)nvecs * c->ndof);
  c->work_n = c->ndof; c->work_vecs = nvecs;
}

int solve(Ctx* c, Matrix& M, int method, int precond, double reduction, int maxit, int fixed_iters, const double* d_b, double* d_z,
          mdb_solve_stats* stats)
{
  if (method == MDB_SOLVER_DIRECT) { MDB_REQUIRE(fixed_iters < 0, MDB_EINVAL, "mdb_solve_fixed: not with MDB_SOLVER_DIRECT"); return direct_solve(c, M, d_b, d_z, stats); }
  MDB_REQUIRE(method == MDB_SOLVER_CG || method == MDB_SOLVER_BICGSTAB, MDB_EINVAL, "unknown solver");
  MDB_REQUIRE(precond == MDB_PRECOND_NONE || precond == MDB_PRECOND_JACOBI || precond == MDB_PRECOND_SSOR || precond == MDB_PRECOND_ILU0 || precond == MDB_PRECOND_MG,
              MDB_EINVAL, "unknown preconditioner");
  // Jacobi (and none) are fused into the vector kernels through the inverse diagonal; SSOR / ILU0 are explicit applications
  // z = M^-1 r (precond.cu) between the same kernels run with a unit "inverse diagonal"
  const bool general = precond == MDB_PRECOND_SSOR || precond == MDB_PRECOND_ILU0 || precond == MDB_PRECOND_MG;
  // on a partition SSOR / ILU0 become block-Jacobi over the ranks: each rank applies the sequential preconditioner of its owned
  // block (precond.cu); iteration counts then depend on the number of ranks
  const int64_t n = c->ndof;
  const int vgrid = (cdiv(n, VEC_THREADS) < 148 * 8) ? cdiv(n, VEC_THREADS) : 148 * 8;
  if (!M.d_dinv) { M.d_dinv = dalloc<double>(n); M.dinv_valid = false; }
  if (!M.dinv_valid || M.dinv_precond != precond) {     // the inverse diagonal survives until the matrix values change (mdb_jacobian / mdb_matrix_zero)
    MDB_LAUNCH(c, k_extract_dinv, cdiv(n, 256), 256, 0, n, M.d_rowptr, M.d_diag, M.d_val, M.d_dinv, precond == MDB_PRECOND_JACOBI ? 1 : 0);
    M.dinv_valid = true; M.dinv_precond = precond;
  }
  const bool multi = c->nranks > 1;
  double* s = c->d_scal;
  double* x = d_z;
  int nb = 0;
  MDB_CUDA(cudaEventRecord(c->ev0, c->stream));
  const int check_every = (fixed_iters >= 0) ? (1 << 30) : (general ? 1 : 16);     // a sweep costs far more than a poll
  int issued = 0;
  double* hp = c->h_pinned;
  auto poll = [&]() {
    MDB_CUDA(cudaMemcpyAsync(hp, s, 24 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
  };
  if (method == MDB_SOLVER_CG) {
    ensure_work(c, general ? 4 : 3);
    double *r = c->d_work, *p = c->d_work + n, *q = c->d_work + 2 * n, *zv = general ? c->d_work + 3 * n : nullptr;
    MDB_CUDA(cudaMemsetAsync(s + S_DONE, 0, sizeof(double), c->stream));
    MDB_LAUNCH(c, k_cg_init, vgrid, VEC_THREADS, 0, n, d_b, M.d_dinv, c->d_owned, x, r, p, c->d_partial);
    reduce_ranks(c, multi, vgrid, 2, S_RHONEW, S_RR);
    auto precondition = [&]() {            // zv = M^-1 r ; rho_new = zv . r
      precond_apply(c, M, precond, c->ssor_sweeps, r, zv);
      MDB_LAUNCH(c, k_dot, vgrid, VEC_THREADS, 0, n, s, zv, r, c->d_owned, c->d_partial);
      reduce_ranks(c, multi, vgrid, 1, S_RHONEW);
    };
    if (general) precondition();
    MDB_LAUNCH(c, k_cg_init_scalars, 1, 1, 0, s, reduction);
    while (issued < maxit) {
      MDB_LAUNCH(c, k_cg_update_p, vgrid, VEC_THREADS, 0, n, s, M.d_dinv, general ? zv : r, p);
      if (multi) spmv_halo<1>(c, M, p, q, p, c->d_partial, &nb, s);
      else launch_spmv<1>(c, M, p, q, p, c->d_partial, &nb, s);
      reduce_ranks(c, multi, nb, 1, S_PQ);
      MDB_LAUNCH(c, k_cg_derive_a, 1, 1, 0, s);
      MDB_LAUNCH(c, k_cg_update_xr, vgrid, VEC_THREADS, 0, n, s, M.d_dinv, p, q, c->d_owned, x, r, c->d_partial);
      reduce_ranks(c, multi, vgrid, 2, S_RHONEW, S_RR);
      if (general) precondition();
      MDB_LAUNCH(c, k_cg_derive_b, 1, 1, 0, s);
      ++issued;
      if (issued % check_every == 0) { poll(); if (hp[S_DONE] != 0.0) break; }
    }
  } else {
    ensure_work(c, 7);
    double *r = c->d_work, *rt = r + n, *p = rt + n, *v = p + n, *y = v + n, *y2 = y + n, *t = y2 + n;
    MDB_LAUNCH(c, k_bi_init, vgrid, VEC_THREADS, 0, n, d_b, c->d_owned, x, r, rt, p, v, c->d_partial);
    reduce_ranks(c, multi, vgrid, 1, S_RR);
    MDB_LAUNCH(c, k_bi_init_scalars, 1, 1, 0, s, reduction);
    while (issued < maxit) {
      MDB_LAUNCH(c, k_bi_update_p, vgrid, VEC_THREADS, 0, n, s, M.d_dinv, r, v, p, y);
      if (general) precond_apply(c, M, precond, c->ssor_sweeps, p, y);
      if (multi) spmv_halo<1>(c, M, y, v, rt, c->d_partial, &nb, s);
      else launch_spmv<1>(c, M, y, v, rt, c->d_partial, &nb, s);
      reduce_ranks(c, multi, nb, 1, S_H);
      MDB_LAUNCH(c, k_bi_derive_alpha, 1, 1, 0, s);
      MDB_LAUNCH(c, k_bi_update_s, vgrid, VEC_THREADS, 0, n, s, M.d_dinv, y, v, c->d_owned, x, r, y2, c->d_partial);
      reduce_ranks(c, multi, vgrid, 1, S_RR);
      MDB_LAUNCH(c, k_bi_derive_half, 1, 1, 0, s);
      if (general) precond_apply(c, M, precond, c->ssor_sweeps, r, y2);
      if (multi) spmv_halo<2>(c, M, y2, t, r, c->d_partial, &nb, s);
      else launch_spmv<2>(c, M, y2, t, r, c->d_partial, &nb, s);
      reduce_ranks(c, multi, nb, 2, S_TR, S_TT);
      MDB_LAUNCH(c, k_bi_derive_omega, 1, 1, 0, s);
      MDB_LAUNCH(c, k_bi_update_x, vgrid, VEC_THREADS, 0, n, s, y2, t, rt, c->d_owned, x, r, c->d_partial);
      reduce_ranks(c, multi, vgrid, 2, S_RR, S_PQ);
      MDB_LAUNCH(c, k_bi_derive_end, 1, 1, 0, s);
      ++issued;
      if (issued % check_every == 0) { poll(); if (hp[S_DONE] != 0.0) break; }
    }
  }
  MDB_CUDA(cudaEventRecord(c->ev1, c->stream));
  poll();
  float ms = 0.f;
  MDB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->phases["solve"].last_ms = ms;
  const bool converged = hp[S_DONE] != 0.0 && hp[S_BREAK] == 0.0;
  if (multi && halo_timed_out(c)) throw Error(MDB_ENCCL, "halo exchange timed out waiting for a neighbour");
  if (stats) {
    stats->iterations = (int)hp[S_ITER];
    stats->converged = converged ? 1 : 0;
    stats->defect0 = std::sqrt(hp[S_RR0]);
    stats->defect = std::sqrt(hp[S_RR]);
    stats->seconds = ms * 1e-3;
  }
  if (fixed_iters >= 0) return MDB_OK;
  if (!converged) { c->err = "linear solver did not converge within maxit iterations"; return MDB_ENOTCONVERGED; }
  return MDB_OK;
}

} // namespace mdb
