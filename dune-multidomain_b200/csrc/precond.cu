// precond.cu — the reference's sequential preconditioners on the device: SeqSSOR and SeqILU0 with ISTL's arithmetic.
//
// Replaces [UPSTREAM] dune-istl SeqSSOR (gsetc.hh bsorf/bsorb) and SeqILU0 (ilu.hh bilu0_decomposition / bilu_backsolve)
// as PDELab's ISTLBackend_SEQ_BCGS_SSOR / ISTLBackend_SEQ_CG_SSOR / *_ILU0 drive them (test/testpoisson.cc:291-292,
// test/explicitlycoupledpoisson.cc:297-313).  Both are Gauss-Seidel-like recurrences over the rows in index order.  They
// are run here WITHOUT a level analysis and WITHOUT changing a single operation of the sequential algorithm:
//   * one warp owns one row; warps take rows from a global counter in a WAVEFRONT order: rows sorted by (child block,
//     ix + 2 iy + 4 iz) of their lattice point, a topological order of the lower-triangular dependency graph of a
//     lexicographically numbered lattice block (every lower stencil neighbour has a strictly smaller key; other blocks come
//     first).  The rows in flight are then (nearly) independent, so the whole wavefront works in parallel; the ARITHMETIC is
//     untouched by the order in which independent rows are executed;
//   * a row that needs the new value of row j spins on ready[j] == epoch (acquire) and then reads v[j] from L2;
//   * the owner publishes v[i] and then ready[i] = epoch (release).  Every row a warp can wait for has already been taken
//     by a running warp (the dispatch order is topological), so the scheme cannot deadlock, whatever the grid size;
//   * the row sum is accumulated in column order, one subtraction at a time, products rounded separately (no FMA) — the
//     bits of the sequential CPU sweep, so Krylov iteration counts match the reference's solver stack.
// The sweeps are latency-bound (critical path = longest dependency chain, ~7N rows on an N^3 lattice); they are the
// drop-in for the reference's default solver configuration, the fast path is Jacobi (solver.cu).
// Assumption: the pattern is structurally symmetric (true for every pattern this backend builds: unions of cliques and
// two-sided coupling blocks), so a row never reads an entry that a later row of the same sweep is already overwriting.
#include "common.cuh"
#include <cub/cub.cuh>

namespace mdb {

__device__ __forceinline__ int ld_acquire_s32(const int* p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v)
{
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

enum { SW_SSOR_FWD = 0, SW_SSOR_BWD = 1, SW_ILU_LOWER = 2, SW_ILU_UPPER = 3 };

// The iterate travels between warps as self-validating 16-byte records (the layout of NCCL's LL protocol): two 64-bit words,
// each = (tag << 32) | half of the double.  64-bit stores are single-copy atomic, so a reader that sees the sweep's tag in
// BOTH words has the complete value — one store and one polling load per dependency, no fence, no separate flag.
__device__ __forceinline__ ulonglong2 ld_record(const ulonglong2* p)
{
  ulonglong2 w;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(p) : "memory");
  return w;
}
__device__ __forceinline__ void st_record(ulonglong2* p, double v, unsigned tag)
{
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long t = (unsigned long long)tag << 32;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(t | (b & 0xffffffffull)), "l"(t | (b >> 32)) : "memory");
}
__device__ __forceinline__ double record_value(const ulonglong2& w) { return __longlong_as_double((long long)((w.y << 32) | (w.x & 0xffffffffull))); }
__device__ __forceinline__ bool record_tagged(const ulonglong2& w, unsigned tag) { return (unsigned)(w.x >> 32) == tag && (unsigned)(w.y >> 32) == tag; }

// One sweep.  a = matrix values (SSOR) or the ILU factors; d = right-hand side; vt = the iterate as records (updated in place).
//   SSOR fwd/bwd (gsetc.hh bsorf/bsorb, w = 1):  rhs = d_i - sum_j a_ij v_j over ALL j of the row (j < i new / j >= i old in
//                                                the forward sweep, mirrored in the backward one); v_i += rhs / a_ii
//   ILU lower (bilu_backsolve, first loop):      v_i = d_i - sum_{j<i} L_ij v_j
//   ILU upper (second loop):                     v_i = Uinv_ii * (v_i - sum_{j>i} U_ij v_j), the diagonal holds the inverse
template <int MODE>
__global__ void __launch_bounds__(256) k_sweep(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                               const int32_t* __restrict__ diag, const double* __restrict__ a, const double* __restrict__ d,
                                               ulonglong2* vt, const int32_t* __restrict__ perm, unsigned epoch, unsigned long long* counter,
                                               const uint8_t* __restrict__ owned)
{
  constexpr bool ASC = (MODE == SW_SSOR_FWD || MODE == SW_ILU_LOWER);
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(counter, 1ull);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= (unsigned long long)n) return;
    const int64_t i = perm[ASC ? (int64_t)k : n - 1 - (int64_t)k];
    // partitioned context: block-Jacobi over the ranks — the sweep runs on the owned x owned block of the local matrix; ghost
    // rows carry 0 and ghost columns are left out
    if (owned && !owned[i]) { if (lane == 0) st_record(vt + i, 0.0, epoch); __syncwarp(); continue; }
    const int64_t s = rowptr[i], e = rowptr[i + 1];
    const bool norow = diag[i] < 0;          // a row without entries (a block no participant of the matrix's operators lives on): identity row
    const int64_t pd = norow ? s : s + diag[i];
    const double aii = norow ? 1.0 : a[pd];                                 // issued early: needed only at the end of the chain
    const double vi = (MODE == SW_ILU_LOWER) ? 0.0 : record_value(ld_record(vt + i));   // row i is written by this warp only
    // the entries this mode sums, in column order
    int64_t p0 = s, p1 = e;
    if (MODE == SW_ILU_LOWER) p1 = pd;
    if (MODE == SW_ILU_UPPER) p0 = norow ? e : pd + 1;
    double acc = (MODE == SW_ILU_UPPER) ? vi : d[i];
    for (int64_t pb = p0; pb < p1; pb += 32) {
      const int64_t p = pb + lane;
      bool live = p < p1;
      const int col = live ? colidx[p] : 0;
      if (live && owned && !owned[col]) live = false;
      const double av = live ? a[p] : 0.0;
      const bool dep = live && (ASC ? (col < i) : (col > i));
      ulonglong2 w = make_ulonglong2(0ull, 0ull);
      for (int tries = 0;; ++tries) {                                       // warp-wide wait: the row needs all its dependencies anyway
        if (live) w = ld_record(vt + col);
        if (__all_sync(0xffffffffu, !dep || record_tagged(w, epoch))) break;
        __nanosleep(tries < 8 ? 32 : 128);
      }
      const double prod = live ? __dmul_rn(av, record_value(w)) : 0.0;
      const int cnt = (p1 - pb < 32) ? (int)(p1 - pb) : 32;
      for (int l = 0; l < cnt; ++l) acc = __dsub_rn(acc, __shfl_sync(0xffffffffu, prod, l));
    }
    if (lane == 0) {
      double out;
      if (norow) out = acc;
      else if (MODE == SW_SSOR_FWD || MODE == SW_SSOR_BWD) out = __dadd_rn(vi, __ddiv_rn(acc, aii));
      else if (MODE == SW_ILU_LOWER) out = acc;
      else out = __dmul_rn(aii, acc);
      st_record(vt + i, out, epoch);
    }
    __syncwarp();
  }
}

__global__ void k_records_to_vector(int64_t n, const ulonglong2* __restrict__ vt, double* v)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = record_value(vt[i]);
}

// ILU(0) on the pattern of A (ilu.hh bilu0_decomposition): for every row i, for every k < i of the row in ascending order:
//   L_ik = a_ik * Uinv_kk ; a_ij -= L_ik * U_kj for the j > k present in both rows ; finally the diagonal is inverted.
// A warp owns row i and keeps its entries in registers (EPL per lane); row k's upper part is streamed through the lanes.
template <int EPL>
__global__ void __launch_bounds__(256) k_ilu0_factor(int64_t n, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                     const int32_t* __restrict__ diag, double* f, const int32_t* __restrict__ perm, int* ready, int epoch,
                                                     unsigned long long* counter, const uint8_t* __restrict__ owned)
{
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long kk = 0;
    if (lane == 0) kk = atomicAdd(counter, 1ull);
    kk = __shfl_sync(0xffffffffu, kk, 0);
    if (kk >= (unsigned long long)n) return;
    const int64_t i = perm[kk];
    if (owned && !owned[i]) { if (lane == 0) st_release_s32(ready + i, epoch); __syncwarp(); continue; }     // ghost row: not part of the block
    const int64_t s = rowptr[i];
    const int len = (int)(rowptr[i + 1] - s), dg = diag[i];
    int col[EPL]; double val[EPL];
#pragma unroll
    for (int u = 0; u < EPL; ++u) {
      const int idx = u * 32 + lane;
      col[u] = (idx < len) ? colidx[s + idx] : -1;
      val[u] = (idx < len) ? f[s + idx] : 0.0;        // f starts as a copy of A; row i is touched by this warp only
    }
    for (int pl = 0; pl < dg; ++pl) {                  // the entries left of the diagonal, ascending
      int csel = col[0];
#pragma unroll
      for (int u = 1; u < EPL; ++u) if ((pl >> 5) == u) csel = col[u];       // pl >> 5 is warp-uniform
      const int k = __shfl_sync(0xffffffffu, csel, pl & 31);
      if (owned && !owned[k]) continue;                                      // a ghost column: outside the owned x owned block (warp-uniform)
      for (int tries = 0; ld_acquire_s32(ready + k) != epoch; ++tries) __nanosleep(tries < 8 ? 32 : 128);
      const int64_t ks = rowptr[k];
      const int kd = diag[k], klen = (int)(rowptr[k + 1] - ks);
      const double uinv = __ldcg(f + ks + kd);
      double lik = 0.0;
#pragma unroll
      for (int u = 0; u < EPL; ++u)
        if (u * 32 + lane == pl) { val[u] = __dmul_rn(val[u], uinv); lik = val[u]; }
      lik = __shfl_sync(0xffffffffu, lik, pl & 31);
      for (int qb = kd + 1; qb < klen; qb += 32) {
        const int q = qb + lane;
        const int cq = (q < klen) ? colidx[ks + q] : -2;
        const double uq = (q < klen) ? __ldcg(f + ks + q) : 0.0;
        const int cnt = (klen - qb < 32) ? klen - qb : 32;
        for (int t = 0; t < cnt; ++t) {
          const int c = __shfl_sync(0xffffffffu, cq, t);
          const double uu = __shfl_sync(0xffffffffu, uq, t);
#pragma unroll
          for (int u = 0; u < EPL; ++u)
            if (col[u] == c) val[u] = __dsub_rn(val[u], __dmul_rn(lik, uu));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < EPL; ++u) {
      const int idx = u * 32 + lane;
      if (idx == dg) val[u] = __ddiv_rn(1.0, val[u]);  // (*ij).invert()
      if (idx < len) __stcg(f + s + idx, val[u]);
    }
    __threadfence();                                   // every lane's stores are visible device-wide ...
    __syncwarp();
    if (lane == 0) st_release_s32(ready + i, epoch);   // ... before the row is published
  }
}

// wavefront key of every row: (child block, ix + 2 iy + 4 iz); QkDG blocks: the coordinates of the row's CELL (st = lattice stride)
struct KeyChild { int64_t offset; int n0, n1; const int32_t* dof2lat; int st; };
struct KeySpace { int nchild; KeyChild ch[MDB_MAX_CHILDREN]; };
__global__ void k_sweep_keys(int64_t n, KeySpace sp, uint32_t* key, int32_t* row)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int ci = 0;
  while (ci + 1 < sp.nchild && r >= sp.ch[ci + 1].offset) ++ci;
  const KeyChild& ch = sp.ch[ci];
  const int64_t p = ch.dof2lat[r - ch.offset];
  const int ix = (int)(p % ch.n0) / ch.st, iy = (int)((p / ch.n0) % ch.n1) / ch.st, iz = (int)(p / ((int64_t)ch.n0 * ch.n1)) / ch.st;
  key[r] = ((uint32_t)ci << 24) | (uint32_t)(ix + 2 * iy + 4 * iz);
  row[r] = (int32_t)r;
}
// ---- QkDG matrices: the sweeps cell by cell -----------------------------------------------------------------------------------------
// The rows of a DG cell form a dense diagonal block, so inside a cell every row depends on the row before it: with one warp per ROW
// the critical path of a sweep is (rows per cell) x (cells on the longest wavefront) round trips through L2 (~1.8 us each).  Here a
// warp owns a CELL: it stages the cell's value image (contiguous: the rows share their column set) and the iterate at the cell's
// columns in shared memory, lane i carries row i, and the in-cell recurrence runs through shared memory — only whole cells talk to
// each other through the record array.  Arithmetic and its order are those of the row kernel (= the sequential sweep): the row sum
// in column order, one separately rounded product and one subtraction at a time, v_i += rhs / a_ii.
//   forward :  row i = [lower blocks, new] [own columns < i, new] [own columns >= i, old] [upper blocks, old]
//   backward:  row i = [lower blocks, old] [own columns <= i, old] [own columns > i, new] [upper blocks, new]
// Cells are dispatched in wavefront order of their coordinates (cx + cy + cz inside a child block, child blocks in order), a
// topological order of the cell dependency graph (face neighbours and earlier child blocks) — the deadlock argument of k_sweep.
struct CellSpace { int nchild; int64_t offset[MDB_MAX_CHILDREN]; int nl[MDB_MAX_CHILDREN]; };

template <bool FWD>
__global__ void __launch_bounds__(256) k_sweep_cells(int64_t ncells, const int32_t* __restrict__ cell_perm, CellSpace sp, int img_doubles, int maxrow,
                                                     const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                     const double* __restrict__ a, const double* __restrict__ d, ulonglong2* vt, unsigned epoch,
                                                     unsigned long long* counter)
{
  extern __shared__ double s_cells[];
  const int lane = threadIdx.x & 31;
  double* aimg = s_cells + (size_t)(threadIdx.x >> 5) * (img_doubles + maxrow);
  double* vv = aimg + img_doubles;
  for (;;) {
    unsigned long long k = 0;
    if (lane == 0) k = atomicAdd(counter, 1ull);
    k = __shfl_sync(0xffffffffu, k, 0);
    if (k >= (unsigned long long)ncells) return;
    const int64_t row0 = cell_perm[FWD ? (int64_t)k : ncells - 1 - (int64_t)k];
    int ci = 0;
    while (ci + 1 < sp.nchild && row0 >= sp.offset[ci + 1]) ++ci;
    const int nl = sp.nl[ci];
    const int64_t rs = rowptr[row0];
    const int len = (int)(rowptr[row0 + 1] - rs);
    if (len == 0) {                                       // rows without entries (a block no participant of the matrix's operators lives on): identity rows
      if (lane < nl) st_record(vt + row0 + lane, d[row0 + lane], epoch);
      __syncwarp();
      continue;
    }
    __syncwarp();
    for (int e = lane; e < nl * len; e += 32) aimg[e] = a[rs + e];
    // the iterate at the cell's columns (at most 64: two per lane).  All records are requested at once, so the old values the cell
    // needs anyway arrive while it waits (warp-wide, as in k_sweep) for the columns of the cells it depends on
    const int c1 = lane + 32;
    const bool live0 = lane < len, live1 = c1 < len;
    const int col0 = live0 ? colidx[rs + lane] : 0, col1 = live1 ? colidx[rs + c1] : 0;
    const bool dep0 = live0 && (FWD ? (col0 < row0) : (col0 >= row0 + nl)), dep1 = live1 && (FWD ? (col1 < row0) : (col1 >= row0 + nl));
    ulonglong2 w0 = make_ulonglong2(0ull, 0ull), w1 = w0;
    if (live0) w0 = ld_record(vt + col0);
    if (live1) w1 = ld_record(vt + col1);
    for (int tries = 0;; ++tries) {
      const bool ok0 = !dep0 || record_tagged(w0, epoch), ok1 = !dep1 || record_tagged(w1, epoch);
      if (__all_sync(0xffffffffu, ok0 && ok1)) break;
      __nanosleep(tries < 8 ? 32 : 128);
      if (!ok0) w0 = ld_record(vt + col0);
      if (!ok1) w1 = ld_record(vt + col1);
    }
    if (live0) vv[lane] = record_value(w0);
    if (live1) vv[c1] = record_value(w1);
    int own_start = 0;
    {
      const unsigned hit0 = __ballot_sync(0xffffffffu, live0 && col0 == (int)row0), hit1 = __ballot_sync(0xffffffffu, live1 && col1 == (int)row0);
      own_start = hit0 ? __ffs(hit0) - 1 : 32 + __ffs(hit1) - 1;
    }
    __syncwarp();
    const bool mine = lane < nl;
    const double* ar = aimg + (size_t)(mine ? lane : 0) * len;
    double acc = mine ? d[row0 + lane] : 0.0;
    if (FWD) {
      if (mine) for (int c = 0; c < own_start; ++c) acc = __dsub_rn(acc, __dmul_rn(ar[c], vv[c]));
      for (int j = 0; j < nl; ++j) {
        if (lane == j) {
          for (int c = own_start + j; c < len; ++c) acc = __dsub_rn(acc, __dmul_rn(ar[c], vv[c]));
          vv[own_start + j] = __dadd_rn(vv[own_start + j], __ddiv_rn(acc, ar[own_start + j]));
        }
        __syncwarp();
        if (mine && lane > j) acc = __dsub_rn(acc, __dmul_rn(ar[own_start + j], vv[own_start + j]));
      }
    } else {
      if (mine) for (int c = 0; c <= own_start + lane; ++c) acc = __dsub_rn(acc, __dmul_rn(ar[c], vv[c]));
      for (int j = nl - 1; j >= 0; --j) {
        if (lane == j) {
          for (int c = own_start + j + 1; c < len; ++c) acc = __dsub_rn(acc, __dmul_rn(ar[c], vv[c]));
          // the diagonal term a_jj v_j (old) was subtracted in the prefix; v_j += rhs / a_jj
          vv[own_start + j] = __dadd_rn(vv[own_start + j], __ddiv_rn(acc, ar[own_start + j]));
        }
        __syncwarp();
      }
    }
    __syncwarp();
    if (mine) st_record(vt + row0 + lane, vv[own_start + lane], epoch);
    __syncwarp();
  }
}

__global__ void k_cell_keys(int64_t ncells, CellSpace sp, KeySpace ks, uint32_t* key, int32_t* row)
{
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ncells) return;
  int ci = 0; int64_t first = 0;                       // cells are enumerated child block after child block
  while (ci + 1 < sp.nchild) { const int64_t nc = (sp.offset[ci + 1] - sp.offset[ci]) / sp.nl[ci]; if (t < first + nc) break; first += nc; ++ci; }
  const int64_t row0 = sp.offset[ci] + (t - first) * sp.nl[ci];
  const KeyChild& ch = ks.ch[ci];
  const int64_t p = ch.dof2lat[row0 - ch.offset];
  const int cx = (int)(p % ch.n0) / ch.st, cy = (int)((p / ch.n0) % ch.n1) / ch.st, cz = (int)(p / ((int64_t)ch.n0 * ch.n1)) / ch.st;
  key[t] = ((uint32_t)ci << 24) | (uint32_t)(cx + cy + cz);
  row[t] = (int32_t)row0;
}

__global__ void k_iota(int64_t n, int32_t* row)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < n) row[r] = (int32_t)r;
}

static void ensure_precond_state(Ctx* c, Matrix& M)
{
  if (M.d_ready) return;
  const int64_t n = c->ndof;
  M.d_ready = dalloc<int>(n);
  MDB_CUDA(cudaMemsetAsync(M.d_ready, 0, n * sizeof(int), c->stream));
  M.d_sweep_counter = dalloc<unsigned long long>(1);
  M.sweep_epoch = 0;
  // dispatch order of the rows.  The wavefront key is a topological order only for lexicographically numbered (Q1) lattice
  // blocks whose lattice fits the key; anything else keeps the index order (always topological, little parallelism).
  // QkDG blocks are numbered cell block after cell block, cells in host order: a row depends on the earlier rows of its own cell
  // (same key; the stable sort keeps them in index order) and on the cells at -x, -y, -z (smaller key) — the same key on cell
  // coordinates is topological.
  M.d_sweep_perm = dalloc<int32_t>(n);
  bool lattice = c->um == nullptr;            // unstructured meshes: index order
  KeySpace sp; sp.nchild = (int)c->children.size();
  for (int i = 0; i < sp.nchild; ++i) {
    const Child& ch = c->children[i];
    lattice = lattice && (ch.k == 1 || ch.dg) && (int64_t)ch.lat_n[0] + 2 * (int64_t)ch.lat_n[1] + 4 * (int64_t)ch.lat_n[2] < (1 << 24);
    sp.ch[i].offset = ch.offset; sp.ch[i].n0 = ch.lat_n[0]; sp.ch[i].n1 = ch.lat_n[1]; sp.ch[i].dof2lat = ch.d_dof2lat; sp.ch[i].st = ch.dg ? ch.stride() : 1;
  }
  // all-DG systems whose cell image fits a warp's share of shared memory: the cell-by-cell sweeps (k_sweep_cells)
  {
    bool all_dg = !c->children.empty() && !c->mesh.partitioned;
    CellSpace cs; cs.nchild = (int)c->children.size();
    int nlmax = 1; int64_t ncells = 0;
    for (int i = 0; i < cs.nchild; ++i) {
      const Child& ch = c->children[i];
      all_dg = all_dg && ch.dg;
      cs.offset[i] = ch.offset; cs.nl[i] = ch.nloc(c->mesh.dim);
      if (cs.nl[i] > nlmax) nlmax = cs.nl[i];
      ncells += ch.size / cs.nl[i];
    }
    static const bool off = getenv("MDB_SWEEP_ROWS") != nullptr;          // A/B switch: the row kernel on DG matrices
    if (all_dg && !off && lattice && ncells > 0 && M.max_row <= 64 && (size_t)(nlmax * M.max_row + M.max_row) * sizeof(double) <= 6 * 1024) {
      M.sweep_cells = ncells; M.sweep_cell_img = nlmax * M.max_row;
      M.d_cell_perm = dalloc<int32_t>(ncells);
      uint32_t* d_key = dalloc<uint32_t>(ncells); uint32_t* d_key2 = dalloc<uint32_t>(ncells);
      int32_t* d_row = dalloc<int32_t>(ncells);
      MDB_LAUNCH(c, k_cell_keys, cdiv(ncells, 256), 256, 0, ncells, cs, sp, d_key, d_row);
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key, d_key2, d_row, M.d_cell_perm, (int)ncells, 0, 32, c->stream);
      void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
      MDB_CUDA(cub::DeviceRadixSort::SortPairs(d_t, tb, d_key, d_key2, d_row, M.d_cell_perm, (int)ncells, 0, 32, c->stream)); c->launches++;
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(d_t); dfree(d_key); dfree(d_key2); dfree(d_row);
    }
  }
  if (!lattice) { MDB_LAUNCH(c, k_iota, cdiv(n, 256), 256, 0, n, M.d_sweep_perm); return; }
  uint32_t* d_key = dalloc<uint32_t>(n); uint32_t* d_key2 = dalloc<uint32_t>(n);
  int32_t* d_row = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_sweep_keys, cdiv(n, 256), 256, 0, n, sp, d_key, d_row);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, d_key, d_key2, d_row, M.d_sweep_perm, (int)n, 0, 32, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  MDB_CUDA(cub::DeviceRadixSort::SortPairs(d_t, tb, d_key, d_key2, d_row, M.d_sweep_perm, (int)n, 0, 32, c->stream)); c->launches++;   // stable: ties keep index order
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_t); dfree(d_key); dfree(d_key2); dfree(d_row);
}

template <int MODE>
static void sweep(Ctx* c, Matrix& M, const double* a, const double* d)
{
  MDB_CUDA(cudaMemsetAsync(M.d_sweep_counter, 0, sizeof(unsigned long long), c->stream));
  ++M.sweep_epoch;
  if (M.sweep_cells > 0 && (MODE == SW_SSOR_FWD || MODE == SW_SSOR_BWD)) {
    CellSpace cs; cs.nchild = (int)c->children.size();
    for (int i = 0; i < cs.nchild; ++i) { cs.offset[i] = c->children[i].offset; cs.nl[i] = c->children[i].nloc(c->mesh.dim); }
    int grid = 148 * 6;
    if (grid > cdiv(M.sweep_cells, 8)) grid = cdiv(M.sweep_cells, 8);
    const size_t smem = (size_t)8 * (M.sweep_cell_img + M.max_row) * sizeof(double);
    MDB_LAUNCH(c, (k_sweep_cells<MODE == SW_SSOR_FWD>), grid, 256, smem, M.sweep_cells, M.d_cell_perm, cs, M.sweep_cell_img, M.max_row, M.d_rowptr, M.d_colidx,
               a, d, M.d_vrec, (unsigned)M.sweep_epoch, M.d_sweep_counter);
    return;
  }
  int grid = 148 * 8;
  if (grid > cdiv(c->ndof, 8)) grid = cdiv(c->ndof, 8);
  MDB_LAUNCH(c, k_sweep<MODE>, grid, 256, 0, c->ndof, M.d_rowptr, M.d_colidx, M.d_diag, a, d, M.d_vrec, M.d_sweep_perm, (unsigned)M.sweep_epoch,
             M.d_sweep_counter, c->d_owned);
}

void ilu0_factorize(Ctx* c, Matrix& M)
{
  ensure_precond_state(c, M);
  if (!M.d_ilu) M.d_ilu = dalloc<double>(M.nnz);
  MDB_CUDA(cudaMemcpyAsync(M.d_ilu, M.d_val, M.nnz * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  MDB_CUDA(cudaMemsetAsync(M.d_sweep_counter, 0, sizeof(unsigned long long), c->stream));
  ++M.sweep_epoch;
  int grid = 148 * 8;
  if (grid > cdiv(c->ndof, 8)) grid = cdiv(c->ndof, 8);
  MDB_REQUIRE(M.max_row <= 256, MDB_EINVAL, "ILU0: rows longer than 256 entries are not supported");
  if (M.max_row <= 32)
    MDB_LAUNCH(c, k_ilu0_factor<1>, grid, 256, 0, c->ndof, M.d_rowptr, M.d_colidx, M.d_diag, M.d_ilu, M.d_sweep_perm, M.d_ready, M.sweep_epoch, M.d_sweep_counter, c->d_owned);
  else if (M.max_row <= 64)
    MDB_LAUNCH(c, k_ilu0_factor<2>, grid, 256, 0, c->ndof, M.d_rowptr, M.d_colidx, M.d_diag, M.d_ilu, M.d_sweep_perm, M.d_ready, M.sweep_epoch, M.d_sweep_counter, c->d_owned);
  else if (M.max_row > 128)
    MDB_LAUNCH(c, k_ilu0_factor<8>, grid, 256, 0, c->ndof, M.d_rowptr, M.d_colidx, M.d_diag, M.d_ilu, M.d_sweep_perm, M.d_ready, M.sweep_epoch, M.d_sweep_counter, c->d_owned);
  else
    MDB_LAUNCH(c, k_ilu0_factor<4>, grid, 256, 0, c->ndof, M.d_rowptr, M.d_colidx, M.d_diag, M.d_ilu, M.d_sweep_perm, M.d_ready, M.sweep_epoch, M.d_sweep_counter, c->d_owned);
  M.ilu_valid = true;
}

// v = M^-1 d  (v starts at zero, as the ISTL solvers do: pre() / apply() with v = 0)
void precond_apply(Ctx* c, Matrix& M, int kind, int sweeps, const double* d, double* v)
{
  if (kind == MDB_PRECOND_MG) { mg_apply(c, M, d, v); return; }      // mg.cu: one V-cycle
  const int64_t n = c->ndof;
  // on a partition: block-Jacobi — every rank applies the sequential preconditioner of its owned x owned block (k_sweep's `owned`)
  ensure_precond_state(c, M);
  if (!M.d_vrec) M.d_vrec = dalloc<ulonglong2>(n);
  if (kind == MDB_PRECOND_SSOR) {
    MDB_CUDA(cudaMemsetAsync(M.d_vrec, 0, n * sizeof(ulonglong2), c->stream));      // value 0.0, tag 0 (epochs start at 1)
    for (int it = 0; it < sweeps; ++it) {
      sweep<SW_SSOR_FWD>(c, M, M.d_val, d);
      sweep<SW_SSOR_BWD>(c, M, M.d_val, d);
    }
  } else {
    if (!M.ilu_valid) ilu0_factorize(c, M);
    sweep<SW_ILU_LOWER>(c, M, M.d_ilu, d);
    sweep<SW_ILU_UPPER>(c, M, M.d_ilu, d);
  }
  MDB_LAUNCH(c, k_records_to_vector, cdiv(n, 256), 256, 0, n, M.d_vrec, v);
}

} // namespace mdb
