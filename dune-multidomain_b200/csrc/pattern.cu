// pattern.cu — K2: CSR sparsity pattern by per-row (segmented) sort + unique.
//
// Replaces fill_pattern (gridoperator.hh:83-87 -> patternassemblerengine.hh:142-187,328-337): the pattern is the union
// over cells of the subproblems' FullVolumePattern cliques and, over coupling faces, of FullCouplingPattern's sn / ns
// blocks (couplingutilities.hh:35-47).  Instead of materialising one (row,col) key per local link (64 per Q1 hex) and
// sorting billions of keys, every row is one segment: a thread enumerates the row's candidate columns from the cells
// incident to the row's lattice point, keeps them as a sorted, duplicate-free list (binary insertion), and the row
// pointer is an exclusive scan of the segment lengths.  Pass 1 counts, pass 2 writes; the column order inside a row is
// ascending, i.e. the BCRS order of the reference ([UPSTREAM] (v)).
#include "common.cuh"
#include <cub/cub.cuh>

namespace mdb {

#define MDB_MAXROW 256

struct PChild { int k, subdomain, nloc; int st, dg; int lat_n[3]; int64_t offset, size; const int32_t* lat2dof; const int32_t* dof2lat;
                int iface, lkind, rkind; uint32_t lmask, rmask; };
struct PatSpec {
  int nchild; PChild ch[MDB_MAX_CHILDREN];
  int nsp; int sp_child[16]; int sp_kind[16]; uint32_t sp_mask[16];
  int sp_skel[16];       // doPatternSkeleton: FullSkeletonPattern links to the face neighbours on which the subproblem also applies
  int ncp; int cp_lchild[8], cp_lkind[8], cp_rchild[8], cp_rkind[8]; uint32_t cp_lmask[8], cp_rmask[8];
  int cp_cchild[8];      // >= 0: EnrichedCoupling with this coupling space (sc/cs/nc/cn/cc links instead of sn/ns)
  int cp_ncomp[8];       // component blocks of a power coupling space (children cp_cchild .. cp_cchild + ncomp - 1)
};

__device__ __forceinline__ bool on_cell(int subdomain, uint8_t sd) { return subdomain < 0 || ((sd >> subdomain) & 1); }

struct RowList {
  int32_t v[MDB_MAXROW];
  int n; bool overflow;
  __device__ void insert(int32_t c)
  {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (v[mid] < c) lo = mid + 1; else hi = mid; }
    if (lo < n && v[lo] == c) return;
    if (n >= MDB_MAXROW) { overflow = true; return; }
    for (int i = n; i > lo; --i) v[i] = v[i - 1];
    v[lo] = c; ++n;
  }
};

__device__ void add_cell_dofs(RowList& L, const MeshDesc& m, const PChild& ch, const int ijk[3])
{
  const int k = ch.k, st = ch.st;
  int kz = (m.dim > 2) ? k : 0, ky = (m.dim > 1) ? k : 0;
  for (int az = 0; az <= kz; ++az)
    for (int ay = 0; ay <= ky; ++ay)
      for (int ax = 0; ax <= k; ++ax) {
        int64_t p = (st * ijk[0] + ax) + (int64_t)ch.lat_n[0] * ((st * ijk[1] + ay) + (int64_t)ch.lat_n[1] * (st * ijk[2] + az));
        L.insert((int32_t)(ch.offset + ch.lat2dof[p]));
      }
}

// the mortar DOFs of face f of cell ijk (CouplingLocalFunctionSpace::bind, couplinglocalfunctionspace.hh:208-287)
__device__ void add_face_dofs(RowList& L, const MeshDesc& m, const PChild& C, const int ijk[3], int f)
{
  const int axis = f / 2, side = f % 2;
  for (int i = 0; i < (1 << (m.dim - 1)); ++i) {
    int v[3] = {ijk[0], ijk[1], ijk[2]}; int t = 0;
    for (int d = 0; d < m.dim; ++d) { if (d == axis) v[d] += side; else { v[d] += (i >> t) & 1; ++t; } }
    L.insert((int32_t)(C.offset + C.lat2dof[v[0] + (int64_t)C.lat_n[0] * (v[1] + (int64_t)C.lat_n[1] * v[2])]));
  }
}
__device__ __forceinline__ bool iface_contains(const PChild& C, uint8_t sd_in, uint8_t sd_out)
{
  return cond_applies(C.lkind, C.lmask, sd_in) && cond_applies(C.rkind, C.rmask, sd_out);     // couplinggridfunctionspace.hh:29-35
}

// row of a coupling space: FullEnrichedCoupling{First,Second}SubProblemPattern's cs / cn links and FullEnrichedCouplingPattern's
// cc links (couplingutilities.hh:219-286) over every intersection that contains the row's vertex and on which the coupling fires
__device__ void enumerate_iface_row(RowList& L, const MeshDesc& m, const uint8_t* __restrict__ sd, const PatSpec& S, int child, int64_t gi)
{
  const PChild& C = S.ch[child];
  L.n = 0; L.overflow = false;
  int64_t r = C.dof2lat[gi];
  int p[3]; p[0] = (int)(r % C.lat_n[0]); r /= C.lat_n[0]; p[1] = (int)(r % C.lat_n[1]); p[2] = (int)(r / C.lat_n[1]);
  int lo_c[3], hi_c[3];
  for (int d = 0; d < 3; ++d) {
    if (d >= m.dim) { lo_c[d] = hi_c[d] = 0; continue; }
    hi_c[d] = (p[d] < m.n[d]) ? p[d] : m.n[d] - 1;
    lo_c[d] = (p[d] > 0) ? p[d] - 1 : 0;
  }
  for (int cz = lo_c[2]; cz <= hi_c[2]; ++cz)
    for (int cy = lo_c[1]; cy <= hi_c[1]; ++cy)
      for (int cx = lo_c[0]; cx <= hi_c[0]; ++cx) {
        const int ijk[3] = {cx, cy, cz};
        const uint8_t s = sd[cx + (int64_t)m.n[0] * (cy + (int64_t)m.n[1] * cz)];
        for (int a = 0; a < m.dim; ++a) {
          const int f = 2 * a + (p[a] - ijk[a]);                      // the face of this cell along axis a that holds the vertex
          int nijk[3] = {cx, cy, cz};
          nijk[a] += (f % 2) ? 1 : -1;
          if (nijk[a] < 0 || nijk[a] >= m.n[a]) continue;
          const uint8_t sn = sd[nijk[0] + (int64_t)m.n[0] * (nijk[1] + (int64_t)m.n[1] * nijk[2])];
          if (!iface_contains(C, s, sn)) continue;
          for (int q = 0; q < S.ncp; ++q) {
            if (S.cp_cchild[q] < 0 || child < S.cp_cchild[q] || child >= S.cp_cchild[q] + S.cp_ncomp[q]) continue;
            if (!(cond_applies(S.cp_lkind[q], S.cp_lmask[q], s) && cond_applies(S.cp_rkind[q], S.cp_rmask[q], sn))) continue;
            if (on_cell(S.ch[S.cp_lchild[q]].subdomain, s)) add_cell_dofs(L, m, S.ch[S.cp_lchild[q]], ijk);      // pattern_cs
            if (on_cell(S.ch[S.cp_rchild[q]].subdomain, sn)) add_cell_dofs(L, m, S.ch[S.cp_rchild[q]], nijk);    // pattern_cn
            for (int k = 0; k < S.cp_ncomp[q]; ++k) add_face_dofs(L, m, S.ch[S.cp_cchild[q] + k], ijk, f);      // pattern_cc: full over the power space
          }
        }
      }
}

__device__ void enumerate_row(RowList& L, const MeshDesc& m, const uint8_t* __restrict__ sd, const PatSpec& S, int child, int64_t gi)
{
  const PChild& ch = S.ch[child];
  if (ch.iface) { enumerate_iface_row(L, m, sd, S, child, gi); return; }
  L.n = 0; L.overflow = false;
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  const int k = ch.k, st = ch.st;
  int lo_c[3], hi_c[3];
  for (int d = 0; d < 3; ++d) {
    if (d >= m.dim) { lo_c[d] = hi_c[d] = 0; continue; }
    int c1 = p[d] / st;
    hi_c[d] = (c1 < m.n[d]) ? c1 : m.n[d] - 1;
    lo_c[d] = (!ch.dg && p[d] % k == 0 && c1 > 0) ? c1 - 1 : c1;      // a QkDG point belongs to exactly one cell
    if (lo_c[d] > hi_c[d]) lo_c[d] = hi_c[d];
  }
  for (int cz = lo_c[2]; cz <= hi_c[2]; ++cz)
    for (int cy = lo_c[1]; cy <= hi_c[1]; ++cy)
      for (int cx = lo_c[0]; cx <= hi_c[0]; ++cx) {
        int ijk[3] = {cx, cy, cz};
        uint8_t s = sd[cx + (int64_t)m.n[0] * (cy + (int64_t)m.n[1] * cz)];
        if (!on_cell(ch.subdomain, s)) continue;
        for (int q = 0; q < S.nsp; ++q)
          if (S.sp_child[q] == child && cond_applies(S.sp_kind[q], S.sp_mask[q], s)) { add_cell_dofs(L, m, ch, ijk); break; }
        // invoke_pattern_skeleton_or_boundary (patternassemblerfunctors.hh:36-66) with [UPSTREAM] FullSkeletonPattern: the face is
        // visited from its higher-index cell, pattern_sn / pattern_ns link both cells' rows with both cells' columns
        for (int q = 0; q < S.nsp; ++q) {
          if (S.sp_child[q] != child || !S.sp_skel[q] || !cond_applies(S.sp_kind[q], S.sp_mask[q], s)) continue;
          for (int f = 0; f < 2 * m.dim; ++f) {
            int nijk[3] = {cx, cy, cz};
            nijk[f / 2] += (f % 2) ? 1 : -1;
            if (nijk[f / 2] < 0 || nijk[f / 2] >= m.n[f / 2]) continue;
            const uint8_t sn = sd[nijk[0] + (int64_t)m.n[0] * (nijk[1] + (int64_t)m.n[1] * nijk[2])];
            if (cond_applies(S.sp_kind[q], S.sp_mask[q], sn) && on_cell(ch.subdomain, sn)) add_cell_dofs(L, m, ch, nijk);
          }
        }
        for (int q = 0; q < S.ncp; ++q) {
          bool as_local = S.cp_lchild[q] == child && cond_applies(S.cp_lkind[q], S.cp_lmask[q], s);
          bool as_remote = S.cp_rchild[q] == child && cond_applies(S.cp_rkind[q], S.cp_rmask[q], s);
          if (!as_local && !as_remote) continue;
          for (int f = 0; f < 2 * m.dim; ++f) {
            int nijk[3] = {cx, cy, cz};
            nijk[f / 2] += (f % 2) ? 1 : -1;
            if (nijk[f / 2] < 0 || nijk[f / 2] >= m.n[f / 2]) continue;
            uint8_t sn = sd[nijk[0] + (int64_t)m.n[0] * (nijk[1] + (int64_t)m.n[1] * nijk[2])];
            if (S.cp_cchild[q] >= 0) {       // enriched coupling: pattern_sc (fired from this cell) / pattern_nc (fired from the neighbour)
              for (int k = 0; k < S.cp_ncomp[q]; ++k) {
                const PChild& C = S.ch[S.cp_cchild[q] + k];
                if (as_local && cond_applies(S.cp_rkind[q], S.cp_rmask[q], sn) && iface_contains(C, s, sn)) add_face_dofs(L, m, C, ijk, f);
                if (as_remote && cond_applies(S.cp_lkind[q], S.cp_lmask[q], sn) && iface_contains(C, sn, s)) add_face_dofs(L, m, C, nijk, f ^ 1);
              }
              continue;
            }
            if (as_local && cond_applies(S.cp_rkind[q], S.cp_rmask[q], sn) && on_cell(S.ch[S.cp_rchild[q]].subdomain, sn))
              add_cell_dofs(L, m, S.ch[S.cp_rchild[q]], nijk);      // pattern_sn: rows of the local side, columns of the remote side
            if (as_remote && cond_applies(S.cp_lkind[q], S.cp_lmask[q], sn) && on_cell(S.ch[S.cp_lchild[q]].subdomain, sn))
              add_cell_dofs(L, m, S.ch[S.cp_lchild[q]], nijk);      // pattern_ns: rows of the remote side, columns of the local side
          }
        }
      }
}

__global__ void __launch_bounds__(128) k_pattern_count(MeshDesc m, const uint8_t* __restrict__ sd, PatSpec S, int child, int64_t* rownnz, int* err)
{
  int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (gi >= S.ch[child].size) return;
  RowList L;
  enumerate_row(L, m, sd, S, child, gi);
  if (L.overflow) atomicExch(err, 1);
  rownnz[S.ch[child].offset + gi] = L.n;
}

__global__ void __launch_bounds__(128) k_pattern_fill(MeshDesc m, const uint8_t* __restrict__ sd, PatSpec S, int child,
                                                      const int64_t* __restrict__ rowptr, int32_t* colidx, uint32_t* rowinfo, int32_t* diag)
{
  int64_t gi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const PChild& ch = S.ch[child];
  if (gi >= ch.size) return;
  RowList L;
  enumerate_row(L, m, sd, S, child, gi);
  const int64_t g = ch.offset + gi;
  const int64_t base = rowptr[g];
  uint32_t mask = 0; int nlow = 0; int dg = -1;
  int64_t r = ch.dof2lat[gi];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  for (int i = 0; i < L.n; ++i) {
    int32_t col = L.v[i];
    colidx[base + i] = col;
    if (col == g) dg = i;
    if (col < ch.offset) ++nlow;
    else if (col < ch.offset + ch.size && ch.k == 1 && !ch.dg) {
      int64_t rr = ch.dof2lat[col - ch.offset];
      int q0 = (int)(rr % ch.lat_n[0]); rr /= ch.lat_n[0]; int q1 = (int)(rr % ch.lat_n[1]); int q2 = (int)(rr / ch.lat_n[1]);
      int bit = (q2 - p[2] + 1) * 9 + (q1 - p[1] + 1) * 3 + (q0 - p[0] + 1);
      mask |= (1u << bit);
    }
  }
  rowinfo[g] = mask | ((uint32_t)(nlow > 31 ? 31 : nlow) << 27);
  diag[g] = dg;
}

// "simple" row: full 3^d own-block stencil and nothing else in the row.  flags = 1 for every other (irregular) row;
// heads = 1 where a run of equal simple/irregular status starts.
__global__ void k_irregular_flags(const uint32_t* __restrict__ rowinfo, const int64_t* __restrict__ rowptr, int64_t n, uint32_t full, int ns,
                                  int32_t* flags, int32_t* heads)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto irregular = [&](int64_t r) { return (rowinfo[r] != full || rowptr[r + 1] - rowptr[r] != ns) ? 1 : 0; };   // nlow bits must be 0 too
  const int f = irregular(i);
  flags[i] = f;
  heads[i] = (i == 0 || irregular(i - 1) != f) ? 1 : 0;
}
__global__ void k_irregular_compact(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int64_t n, int32_t* list)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && flags[i]) list[scan[i]] = (int32_t)i;
}

// value-array span [p0, p1) of every run of simple rows (p0 == p1 for a run of irregular rows): what k_jacobian_fill_q1 streams
__global__ void k_seg_spans(const int32_t* __restrict__ seg_start, int64_t nseg, int64_t child_offset, const int64_t* __restrict__ rowptr,
                            const uint32_t* __restrict__ rowinfo, uint32_t full, int ns, longlong2* span)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nseg) return;
  const int64_t g0 = child_offset + seg_start[i], g1 = child_offset + seg_start[i + 1];
  const int64_t p0 = rowptr[g0];
  const bool simple = rowinfo[g0] == full && rowptr[g0 + 1] - p0 == ns;
  span[i] = make_longlong2(p0, simple ? rowptr[g1] : p0);
}

void build_pattern(Ctx* c, Matrix& M)
{
  const MeshDesc& m = c->mesh;
  PatSpec S; std::memset(&S, 0, sizeof(S));
  S.nchild = (int)c->children.size();
  for (int i = 0; i < S.nchild; ++i) {
    const Child& ch = c->children[i];
    PChild& p = S.ch[i];
    p.k = ch.k; p.subdomain = ch.subdomain; p.nloc = ch.nloc(m.dim); p.st = ch.stride(); p.dg = ch.dg ? 1 : 0;
    for (int d = 0; d < 3; ++d) p.lat_n[d] = ch.lat_n[d];
    p.offset = ch.offset; p.size = ch.size; p.lat2dof = ch.d_lat2dof; p.dof2lat = ch.d_dof2lat;
    p.iface = ch.iface ? 1 : 0; p.lkind = ch.lkind; p.lmask = ch.lmask; p.rkind = ch.rkind; p.rmask = ch.rmask;
  }
  for (int gi : M.gos) {
    const GridOp& go = c->gos[gi];
    for (int s : go.sps) {
      const SubProblemDesc& sp = c->sps[s];
      if (!sp.doPatternVolume()) continue;
      MDB_REQUIRE(S.nsp < 16, MDB_EINVAL, "too many subproblems in one matrix");
      S.sp_child[S.nsp] = sp.children[0]; S.sp_kind[S.nsp] = sp.cond_kind; S.sp_mask[S.nsp] = sp.mask; S.sp_skel[S.nsp] = sp.doPatternSkeleton() ? 1 : 0; ++S.nsp;
    }
    for (int k : go.cps) {
      const CouplingDesc& cp = c->cps[k];
      if (cp.op_id == MDB_OP_ND_COUPLING) continue;      // no pattern of its own (test/neumanndirichletcoupling.hh:51-52)
      MDB_REQUIRE(S.ncp < 8, MDB_EINVAL, "too many couplings in one matrix");
      const SubProblemDesc& l = c->sps[cp.local_sp]; const SubProblemDesc& r = c->sps[cp.remote_sp];
      S.cp_lchild[S.ncp] = l.children[0]; S.cp_lkind[S.ncp] = l.cond_kind; S.cp_lmask[S.ncp] = l.mask;
      S.cp_rchild[S.ncp] = r.children[0]; S.cp_rkind[S.ncp] = r.cond_kind; S.cp_rmask[S.ncp] = r.mask;
      S.cp_cchild[S.ncp] = cp.coupling_child; S.cp_ncomp[S.ncp] = cp.ncomp; ++S.ncp;
    }
  }
  const int64_t n = c->ndof;
  int64_t* d_rownnz = dalloc<int64_t>(n + 1);
  M.d_rowptr = dalloc<int64_t>(n + 1);
  int* d_err = dalloc<int>(1);
  MDB_CUDA(cudaMemsetAsync(d_rownnz, 0, (n + 1) * sizeof(int64_t), c->stream));
  MDB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
  for (int i = 0; i < S.nchild; ++i)
    MDB_LAUNCH(c, k_pattern_count, cdiv(S.ch[i].size, 128), 128, 0, m, c->d_cell_sd, S, i, d_rownnz, d_err);
  size_t tmp_bytes = 0, tmp2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_rownnz, M.d_rowptr, (int)(n + 1), c->stream);
  int64_t* d_max = dalloc<int64_t>(1);
  cub::DeviceReduce::Max(nullptr, tmp2, d_rownnz, d_max, (int)n, c->stream);
  if (tmp2 > tmp_bytes) tmp_bytes = tmp2;
  void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_rownnz, M.d_rowptr, (int)(n + 1), c->stream)); c->launches++;
  MDB_CUDA(cub::DeviceReduce::Max(d_tmp, tmp_bytes, d_rownnz, d_max, (int)n, c->stream)); c->launches++;
  int64_t nnz = 0, maxrow = 0; int err = 0;
  MDB_CUDA(cudaMemcpyAsync(&nnz, M.d_rowptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&maxrow, d_max, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp); dfree(d_rownnz); dfree(d_max); dfree(d_err);
  MDB_REQUIRE(!err, MDB_EINVAL, "a matrix row has more than 256 entries (unsupported)");
  M.nnz = nnz; M.max_row = (int)maxrow;
  M.d_colidx = dalloc<int32_t>(nnz);
  M.d_val = dalloc<double>(nnz + 16);     // + 16: the SpMV stages 128-byte aligned windows (solver.cu)
  M.d_rowinfo = dalloc<uint32_t>(n);
  M.d_diag = dalloc<int32_t>(n);
  MDB_CUDA(cudaMemsetAsync(M.d_val, 0, nnz * sizeof(double), c->stream));
  for (int i = 0; i < S.nchild; ++i)
    MDB_LAUNCH(c, k_pattern_fill, cdiv(S.ch[i].size, 128), 128, 0, m, c->d_cell_sd, S, i, M.d_rowptr, M.d_colidx, M.d_rowinfo, M.d_diag);
  // per child: span of its rows in the value array and the compact (ascending) list of irregular rows
  const uint32_t full = (m.dim == 3) ? 0x7FFFFFFu : (0x1FFu << 9);
  int* d_cnt = dalloc<int>(1);
  for (int i = 0; i < S.nchild; ++i) {
    const int64_t sz = S.ch[i].size, off = S.ch[i].offset;
    MDB_CUDA(cudaMemcpyAsync(&M.child_begin[i], M.d_rowptr + off, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(&M.child_end[i], M.d_rowptr + off + sz, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    M.n_irregular[i] = 0; M.d_irregular[i] = nullptr;
    if (sz == 0) continue;
    int32_t* d_flag = dalloc<int32_t>(sz);
    int32_t* d_head = dalloc<int32_t>(sz);
    int32_t* d_scan = dalloc<int32_t>(sz);
    MDB_LAUNCH(c, k_irregular_flags, cdiv(sz, 256), 256, 0, M.d_rowinfo + off, M.d_rowptr + off, sz, full, (m.dim == 3) ? 27 : 9, d_flag, d_head);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_scan, (int)sz, c->stream);
    void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
    MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_flag, d_scan, (int)sz, c->stream)); c->launches++;
    int32_t ls = 0, lf = 0;
    MDB_CUDA(cudaMemcpyAsync(&ls, d_scan + sz - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(&lf, d_flag + sz - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    M.n_irregular[i] = (int64_t)ls + lf;
    M.d_irregular[i] = dalloc<int32_t>(M.n_irregular[i]);
    MDB_LAUNCH(c, k_irregular_compact, cdiv(sz, 256), 256, 0, d_flag, d_scan, sz, M.d_irregular[i]);
    // runs of consecutive rows with the same status: seg_start[0..nseg], seg_start[nseg] = sz
    MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_head, d_scan, (int)sz, c->stream)); c->launches++;
    MDB_CUDA(cudaMemcpyAsync(&ls, d_scan + sz - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(&lf, d_head + sz - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    M.n_seg[i] = (int64_t)ls + lf;
    M.d_seg_start[i] = dalloc<int32_t>(M.n_seg[i] + 1);
    MDB_LAUNCH(c, k_irregular_compact, cdiv(sz, 256), 256, 0, d_head, d_scan, sz, M.d_seg_start[i]);
    const int32_t sz32 = (int32_t)sz;
    MDB_CUDA(cudaMemcpyAsync(M.d_seg_start[i] + M.n_seg[i], &sz32, 4, cudaMemcpyHostToDevice, c->stream));
    M.d_seg_span[i] = dalloc<longlong2>(M.n_seg[i]);
    MDB_LAUNCH(c, k_seg_spans, cdiv(M.n_seg[i], 256), 256, 0, M.d_seg_start[i], M.n_seg[i], off, M.d_rowptr, M.d_rowinfo, full, (m.dim == 3) ? 27 : 9,
               M.d_seg_span[i]);
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t); dfree(d_flag); dfree(d_head); dfree(d_scan);
  }
  dfree(d_cnt);
  MDB_CUDA(cudaStreamSynchronize(c->stream));
}

} // namespace mdb
