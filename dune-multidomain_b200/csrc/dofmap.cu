// dofmap.cu — one-time device DOF map (K1), Dirichlet constraints, interpolation, interface face lists.
//
// Replaces (reference paths): multidomaingridfunctionspace.hh:308-369 (ordering update),
// multidomainlocalfunctionspace.hh:348-358 (bind), constraints.hh:539-558 + constraintsfunctors.hh:41-131,
// interpolate.hh:35-81, coupling.hh:78-83 (which faces a coupling fires on).
#include "common.cuh"
#include <cub/cub.cuh>

namespace mdb {

struct ChildDev {
  int k, subdomain, nloc;
  int st, dg;              // lattice stride (k, or k + 1 for a QkDG child)
  int lat_n[3];
  int64_t offset;
  const int32_t* lat2dof;
  const int32_t* dof2lat;
};

static ChildDev child_dev(const Ctx* c, int i)
{
  const Child& ch = c->children[i];
  ChildDev d; d.k = ch.k; d.subdomain = ch.subdomain; d.nloc = ch.nloc(c->mesh.dim); d.st = ch.stride(); d.dg = ch.dg ? 1 : 0;
  for (int a = 0; a < 3; ++a) d.lat_n[a] = ch.lat_n[a];
  d.offset = ch.offset; d.lat2dof = ch.d_lat2dof; d.dof2lat = ch.d_dof2lat;
  return d;
}

__device__ __forceinline__ void cell_ijk(const MeshDesc& m, int64_t c, int ijk[3])
{
  ijk[0] = (int)(c % m.n[0]); c /= m.n[0]; ijk[1] = (int)(c % m.n[1]); ijk[2] = (int)(c / m.n[1]);
}
__device__ __forceinline__ int64_t cell_index(const MeshDesc& m, const int ijk[3])
{
  return ijk[0] + (int64_t)m.n[0] * (ijk[1] + (int64_t)m.n[1] * ijk[2]);
}
__device__ __forceinline__ bool child_on_cell(int subdomain, uint8_t sd) { return subdomain < 0 || ((sd >> subdomain) & 1); }

// ---- K1a: mark the lattice points in the closure of the child's grid view --------------------------
__global__ void k_mark_present(MeshDesc m, const uint8_t* __restrict__ sd, int k, int st, int subdomain, int lat_n0, int lat_n1, uint8_t* present)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells()) return;
  if (!child_on_cell(subdomain, sd[c])) return;
  int ijk[3]; cell_ijk(m, c, ijk);
  int kz = (m.dim > 2) ? k : 0, ky = (m.dim > 1) ? k : 0;
  for (int az = 0; az <= kz; ++az)
    for (int ay = 0; ay <= ky; ++ay)
      for (int ax = 0; ax <= k; ++ax) {
        int64_t p = (st * ijk[0] + ax) + (int64_t)lat_n0 * ((st * ijk[1] + ay) + (int64_t)lat_n1 * (st * ijk[2] + az));
        present[p] = 1;
      }
}

// ---- K1b: flags of one entity class (set of axes along which the entity extends) --------------------
__global__ void k_class_flags(int dim, int k, int lat_n0, int lat_n1, int64_t nlat, int ext_class, const uint8_t* __restrict__ present,
                              int32_t* flags)
{
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nlat) return;
  int f = 0;
  if (present[p]) {
    int64_t r = p; int e = 0;
    int p0 = (int)(r % lat_n0); r /= lat_n0; int p1 = (int)(r % lat_n1); int p2 = (int)(r / lat_n1);
    if (p0 % k) e |= 1;
    if (dim > 1 && (p1 % k)) e |= 2;
    if (dim > 2 && (p2 % k)) e |= 4;
    f = (e == ext_class);
  }
  flags[p] = f;
}

__global__ void k_assign(int64_t nlat, const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int32_t base,
                         int32_t* lat2dof, int32_t* dof2lat)
{
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nlat) return;
  if (flags[p]) { int32_t g = base + scan[p]; lat2dof[p] = g; dof2lat[g] = (int32_t)p; }
}

__global__ void k_fill_i32(int32_t* a, int64_t n, int32_t v)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

// bounding box of the present lattice points: bb = {min0,min1,min2,max0,max1,max2}
__global__ void k_lattice_bbox(const uint8_t* __restrict__ present, int64_t nlat, int lat_n0, int lat_n1, int* bb)
{
  int mn[3] = {1 << 30, 1 << 30, 1 << 30}, mx[3] = {-1, -1, -1};
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nlat; p += (int64_t)gridDim.x * blockDim.x) {
    if (!present[p]) continue;
    int64_t r = p;
    const int q[3] = {(int)(r % lat_n0), (int)((r / lat_n0) % lat_n1), (int)(r / ((int64_t)lat_n0 * lat_n1))};
    for (int d = 0; d < 3; ++d) { mn[d] = min(mn[d], q[d]); mx[d] = max(mx[d], q[d]); }
  }
  for (int d = 0; d < 3; ++d) {
    for (int o = 16; o > 0; o >>= 1) { mn[d] = min(mn[d], __shfl_down_sync(0xffffffffu, mn[d], o)); mx[d] = max(mx[d], __shfl_down_sync(0xffffffffu, mx[d], o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(bb + d, mn[d]); atomicMax(bb + 3 + d, mx[d]); }
  }
}

// ---- K1a': coupling space — the vertices of the intersections its predicate accepts (couplinggfsordering.hh:113-201).
// The ordering numbers the used vertices by ascending MultiDomainGrid vertex index (entity_dof_offsets is indexed by it and
// partial-summed, :185-200), which is what the class-0 scan below does with these marks.
__global__ void k_mark_present_iface(MeshDesc m, const uint8_t* __restrict__ sd, int lk, uint32_t lmask, int rk, uint32_t rmask, int lat_n0, int lat_n1,
                                     uint8_t* present)
{
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nf = 2 * m.dim;
  if (t >= m.ncells() * nf) return;
  int64_t c = t / nf; int f = (int)(t % nf);
  int ijk[3]; cell_ijk(m, c, ijk);
  const int axis = f / 2, side = f % 2;
  int nijk[3] = {ijk[0], ijk[1], ijk[2]};
  nijk[axis] += side ? 1 : -1;
  if (nijk[axis] < 0 || nijk[axis] >= m.n[axis]) return;                          // !is.neighbor()
  if (!(cond_applies(lk, lmask, sd[c]) && cond_applies(rk, rmask, sd[cell_index(m, nijk)]))) return;
  for (int i = 0; i < (1 << (m.dim - 1)); ++i) {
    int v[3] = {ijk[0], ijk[1], ijk[2]}; int tt = 0;
    for (int d = 0; d < m.dim; ++d) { if (d == axis) v[d] += side; else { v[d] += (i >> tt) & 1; ++tt; } }
    present[v[0] + (int64_t)lat_n0 * (v[1] + (int64_t)lat_n1 * v[2])] = 1;
  }
}

void mark_iface_present(Ctx* c, const Child& ch, uint8_t* d_present)
{
  const MeshDesc& m = c->mesh;
  MDB_LAUNCH(c, k_mark_present_iface, cdiv(m.ncells() * 2 * m.dim, 256), 256, 0, m, c->d_cell_sd, ch.lkind, ch.lmask, ch.rkind, ch.rmask, ch.lat_n[0],
             ch.lat_n[1], d_present);
}

// QkDG numbering ([UPSTREAM] (ii),(iii): all coefficients on the cell): flags of the cells in the child's subdomain, scanned in host
// order, give the subdomain cell index b; DOF = b * nloc + local index
__global__ void k_dg_cell_flags(MeshDesc m, const uint8_t* __restrict__ sd, int subdomain, int32_t* flags)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c < m.ncells()) flags[c] = child_on_cell(subdomain, sd[c]) ? 1 : 0;
}
__global__ void k_dg_assign(MeshDesc m, const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int k, int lat_n0, int lat_n1,
                            int32_t* lat2dof, int32_t* dof2lat)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells() || !flags[c]) return;
  int ijk[3]; cell_ijk(m, c, ijk);
  const int st = k + 1;
  int nl = 1; for (int d = 0; d < m.dim; ++d) nl *= st;
  int kz = (m.dim > 2) ? k : 0, ky = (m.dim > 1) ? k : 0, i = 0;
  for (int az = 0; az <= kz; ++az)
    for (int ay = 0; ay <= ky; ++ay)
      for (int ax = 0; ax <= k; ++ax, ++i) {
        int64_t p = (st * ijk[0] + ax) + (int64_t)lat_n0 * ((st * ijk[1] + ay) + (int64_t)lat_n1 * (st * ijk[2] + az));
        const int32_t g = scan[c] * nl + i;
        lat2dof[p] = g; dof2lat[g] = (int32_t)p;
      }
}

void build_dofmap(Ctx* c)
{
  const MeshDesc& m = c->mesh;
  int64_t off = 0;
  for (auto& ch : c->children) {
    dfree(ch.d_lat2dof); dfree(ch.d_dof2lat);
    ch.nlat = 1;
    for (int d = 0; d < 3; ++d) { ch.lat_n[d] = (d < m.dim) ? (ch.dg ? (ch.k + 1) * m.n[d] : ch.k * m.n[d] + 1) : 1; ch.nlat *= ch.lat_n[d]; }
    MDB_REQUIRE(ch.nlat < (int64_t)2147000000, MDB_EINVAL, "child space too large for int32 device indices");
    uint8_t* d_present = dalloc<uint8_t>(ch.nlat);
    int32_t* d_flags = dalloc<int32_t>(ch.nlat);
    int32_t* d_scan = dalloc<int32_t>(ch.nlat);
    ch.d_lat2dof = dalloc<int32_t>(ch.nlat);
    int32_t* d_dof2lat_tmp = dalloc<int32_t>(ch.nlat);
    MDB_CUDA(cudaMemsetAsync(d_present, 0, ch.nlat, c->stream));
    if (ch.iface) mark_iface_present(c, ch, d_present);
    else MDB_LAUNCH(c, k_mark_present, cdiv(m.ncells(), 256), 256, 0, m, c->d_cell_sd, ch.k, ch.stride(), ch.subdomain, ch.lat_n[0], ch.lat_n[1], d_present);
    MDB_LAUNCH(c, k_fill_i32, cdiv(ch.nlat, 256), 256, 0, ch.d_lat2dof, ch.nlat, -1);
    {
      int h_bb[6] = {1 << 30, 1 << 30, 1 << 30, -1, -1, -1};
      int* d_bb = dalloc<int>(6);
      MDB_CUDA(cudaMemcpyAsync(d_bb, h_bb, sizeof(h_bb), cudaMemcpyHostToDevice, c->stream));
      MDB_LAUNCH(c, k_lattice_bbox, 148 * 4, 256, 0, d_present, ch.nlat, ch.lat_n[0], ch.lat_n[1], d_bb);
      MDB_CUDA(cudaMemcpyAsync(h_bb, d_bb, sizeof(h_bb), cudaMemcpyDeviceToHost, c->stream));
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      dfree(d_bb);
      for (int d = 0; d < 3; ++d) { ch.lat_lo[d] = h_bb[3 + d] < 0 ? 0 : h_bb[d]; ch.lat_hi[d] = h_bb[3 + d] < 0 ? -1 : h_bb[3 + d]; }
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flags, d_scan, (int)ch.nlat, c->stream);
    void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
    int32_t base = 0;
    if (ch.dg) {
      const int64_t nc = m.ncells();
      MDB_REQUIRE(nc <= ch.nlat, MDB_EINVAL, "internal: DG scratch");
      MDB_LAUNCH(c, k_dg_cell_flags, cdiv(nc, 256), 256, 0, m, c->d_cell_sd, ch.subdomain, d_flags);
      MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_scan, (int)nc, c->stream)); c->launches++;
      MDB_LAUNCH(c, k_dg_assign, cdiv(nc, 256), 256, 0, m, d_flags, d_scan, ch.k, ch.lat_n[0], ch.lat_n[1], ch.d_lat2dof, d_dof2lat_tmp);
      int32_t last_scan = 0, last_flag = 0;
      MDB_CUDA(cudaMemcpyAsync(&last_scan, d_scan + nc - 1, 4, cudaMemcpyDeviceToHost, c->stream));
      MDB_CUDA(cudaMemcpyAsync(&last_flag, d_flags + nc - 1, 4, cudaMemcpyDeviceToHost, c->stream));
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      base = (last_scan + last_flag) * ch.nloc(m.dim);
    }
    // [UPSTREAM] ordering: entity dimension ascending (vertex < line < quad < hex); inside one dimension grouped by the
    // set of extension axes (ascending bit pattern); lexicographic inside a group.  For Q1 only class 0 exists.
    for (int mdim = 0; mdim <= m.dim; ++mdim)
      for (int ext = 0; ext < (1 << m.dim); ++ext) {
        if (__builtin_popcount(ext) != mdim) continue;
        if (ch.k == 1 && ext != 0) continue;
        if (ch.dg) continue;
        MDB_LAUNCH(c, k_class_flags, cdiv(ch.nlat, 256), 256, 0, m.dim, ch.k, ch.lat_n[0], ch.lat_n[1], ch.nlat, ext, d_present, d_flags);
        MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_flags, d_scan, (int)ch.nlat, c->stream)); c->launches++;
        MDB_LAUNCH(c, k_assign, cdiv(ch.nlat, 256), 256, 0, ch.nlat, d_flags, d_scan, base, ch.d_lat2dof, d_dof2lat_tmp);
        int32_t last_scan = 0, last_flag = 0;
        MDB_CUDA(cudaMemcpyAsync(&last_scan, d_scan + ch.nlat - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        MDB_CUDA(cudaMemcpyAsync(&last_flag, d_flags + ch.nlat - 1, 4, cudaMemcpyDeviceToHost, c->stream));
        MDB_CUDA(cudaStreamSynchronize(c->stream));
        base += last_scan + last_flag;
      }
    ch.size = base;
    ch.offset = off;
    off += ch.size;
    ch.d_dof2lat = dalloc<int32_t>(ch.size);
    MDB_CUDA(cudaMemcpyAsync(ch.d_dof2lat, d_dof2lat_tmp, ch.size * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_tmp); dfree(d_present); dfree(d_flags); dfree(d_scan); dfree(d_dof2lat_tmp);
  }
  MDB_REQUIRE(off < (int64_t)2147000000, MDB_EINVAL, "more than 2^31 DOFs per context");
  c->ndof = off;
  dfree(c->d_constrained); dfree(c->d_conlist);
  c->d_constrained = dalloc<uint8_t>(c->ndof);
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  c->ncon = 0;
  c->d_conlist = dalloc<int32_t>(1);
}

// ---- parity export of the per-cell local function space -------------------------------------------
struct ChildSet { int n; ChildDev ch[MDB_MAX_CHILDREN]; };

__global__ void k_cell_counts(MeshDesc m, const uint8_t* __restrict__ sd, ChildSet cs, int64_t* counts)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells()) return;
  int n = 0;
  for (int i = 0; i < cs.n; ++i) if (child_on_cell(cs.ch[i].subdomain, sd[c])) n += cs.ch[i].nloc;
  counts[c] = n;
}

__global__ void k_cell_dofs(MeshDesc m, const uint8_t* __restrict__ sd, ChildSet cs, const int64_t* __restrict__ ptr, int64_t* dofs)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells()) return;
  int ijk[3]; cell_ijk(m, c, ijk);
  int64_t o = ptr[c];
  for (int i = 0; i < cs.n; ++i) {
    const ChildDev& ch = cs.ch[i];
    if (!child_on_cell(ch.subdomain, sd[c])) continue;
    int k = ch.k; const int st = ch.st;
    int kz = (m.dim > 2) ? k : 0, ky = (m.dim > 1) ? k : 0;
    for (int az = 0; az <= kz; ++az)
      for (int ay = 0; ay <= ky; ++ay)
        for (int ax = 0; ax <= k; ++ax) {
          int64_t p = (st * ijk[0] + ax) + (int64_t)ch.lat_n[0] * ((st * ijk[1] + ay) + (int64_t)ch.lat_n[1] * (st * ijk[2] + az));
          dofs[o++] = ch.offset + ch.lat2dof[p];
        }
  }
}

static ChildSet make_childset(const Ctx* c)
{
  ChildSet cs; cs.n = (int)c->children.size();
  for (int i = 0; i < cs.n; ++i) cs.ch[i] = child_dev(c, i);
  return cs;
}

void export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs)
{
  const MeshDesc& m = c->mesh;
  const int64_t nc = m.ncells();
  ChildSet cs = make_childset(c);
  int64_t* d_counts = dalloc<int64_t>(nc + 1);
  int64_t* d_ptr = dalloc<int64_t>(nc + 1);
  MDB_CUDA(cudaMemsetAsync(d_counts, 0, (nc + 1) * sizeof(int64_t), c->stream));
  MDB_LAUNCH(c, k_cell_counts, cdiv(nc, 256), 256, 0, m, c->d_cell_sd, cs, d_counts);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, d_ptr, (int)(nc + 1), c->stream);
  void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, d_ptr, (int)(nc + 1), c->stream)); c->launches++;
  MDB_CUDA(cudaMemcpyAsync(cell_ptr, d_ptr, (nc + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  if (cell_dofs) {
    int64_t total = cell_ptr[nc];
    int64_t* d_dofs = dalloc<int64_t>(total);
    MDB_LAUNCH(c, k_cell_dofs, cdiv(nc, 256), 256, 0, m, c->d_cell_sd, cs, d_ptr, d_dofs);
    MDB_CUDA(cudaMemcpyAsync(cell_dofs, d_dofs, total * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    dfree(d_dofs);
  }
  cudaFree(d_tmp); dfree(d_counts); dfree(d_ptr);
}

// ---- constraints --------------------------------------------------------------------------------------
struct ConSpec { int n; int cond_kind[8]; uint32_t mask[8]; int child[8]; mdb_bctype bc[8]; };

// One thread per (cell, face).  constraintsfunctors.hh:77-103: if the subproblem also applies to the outside cell the
// (empty, conforming) skeleton visitor runs; otherwise — domain boundary or subdomain interface — the boundary visitor.
// [UPSTREAM] ConformingDirichletConstraints::boundary: predicate at the face centre; every DOF on the face is constrained.
__global__ void k_constraints(MeshDesc m, const uint8_t* __restrict__ sd, ChildSet cs, ConSpec spec, uint8_t* constrained)
{
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nf = 2 * m.dim;
  if (t >= m.ncells() * nf) return;
  int64_t c = t / nf; int f = (int)(t % nf);
  int axis = f / 2, side = f % 2;
  int ijk[3]; cell_ijk(m, c, ijk);
  uint8_t sde = sd[c];
  int nijk[3] = {ijk[0], ijk[1], ijk[2]};
  nijk[axis] += side ? 1 : -1;
  bool local_nb = nijk[axis] >= 0 && nijk[axis] < m.n[axis];
  bool global_nb = (nijk[axis] + m.off[axis]) >= 0 && (nijk[axis] + m.off[axis]) < m.gn[axis];
  if (global_nb && !local_nb) return;          // partition cut: an interior face whose rows this rank does not own
  uint8_t sdn = local_nb ? sd[cell_index(m, nijk)] : 0;
  // face centre in global coordinates
  double xg[3];
  for (int d = 0; d < m.dim; ++d) {
    double lo = m.coord(d, ijk[d]), up = m.coord(d, ijk[d] + 1);
    xg[d] = (d == axis) ? (side ? up : lo) : lo + 0.5 * (up - lo);
  }
  for (int q = 0; q < spec.n; ++q) {
    if (!cond_applies(spec.cond_kind[q], spec.mask[q], sde)) continue;
    if (local_nb && cond_applies(spec.cond_kind[q], spec.mask[q], sdn)) continue;
    if (eval_bctype(spec.bc[q], m.dim, xg, !global_nb) != 0) continue;
    const ChildDev& ch = cs.ch[spec.child[q]];
    if (!child_on_cell(ch.subdomain, sde)) continue;
    if (ch.dg) continue;           // [UPSTREAM] ConformingDirichletConstraints skips coefficients attached to the cell (codim 0)
    int k = ch.k;
    int kz = (m.dim > 2) ? k : 0, ky = (m.dim > 1) ? k : 0;
    for (int az = 0; az <= kz; ++az)
      for (int ay = 0; ay <= ky; ++ay)
        for (int ax = 0; ax <= k; ++ax) {
          int a[3] = {ax, ay, az};
          if (a[axis] != (side ? k : 0)) continue;
          int64_t p = (k * ijk[0] + ax) + (int64_t)ch.lat_n[0] * ((k * ijk[1] + ay) + (int64_t)ch.lat_n[1] * (k * ijk[2] + az));
          constrained[ch.offset + ch.lat2dof[p]] = 1;
        }
  }
}

void assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n)
{
  MDB_REQUIRE(n <= 8, MDB_EINVAL, "at most 8 constrained subproblems");
  ConSpec spec; spec.n = n;
  for (int q = 0; q < n; ++q) {
    const SubProblemDesc& sp = c->sps[sps[q]];
    spec.cond_kind[q] = sp.cond_kind; spec.mask[q] = sp.mask; spec.child[q] = sp.children[0]; spec.bc[q] = bcs[q];
  }
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  const MeshDesc& m = c->mesh;
  MDB_LAUNCH(c, k_constraints, cdiv(m.ncells() * 2 * m.dim, 256), 256, 0, m, c->d_cell_sd, make_childset(c), spec, c->d_constrained);
}

__global__ void k_u8_to_i32(const uint8_t* in, int32_t* out, int64_t n)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_compact(const uint8_t* __restrict__ flags, const int32_t* __restrict__ scan, int32_t* list, int64_t n)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && flags[i]) list[scan[i]] = (int32_t)i;
}

void rebuild_conlist(Ctx* c)
{
  const int64_t n = c->ndof;
  int32_t* d_f = dalloc<int32_t>(n);
  int32_t* d_s = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_u8_to_i32, cdiv(n, 256), 256, 0, c->d_constrained, d_f, n);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_f, d_s, (int)n, c->stream);
  void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_f, d_s, (int)n, c->stream)); c->launches++;
  int32_t ls = 0, lf = 0;
  MDB_CUDA(cudaMemcpyAsync(&ls, d_s + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&lf, d_f + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->ncon = (int64_t)ls + lf;
  c->con_version++;
  dfree(c->d_conlist);
  c->d_conlist = dalloc<int32_t>(c->ncon);
  MDB_LAUNCH(c, k_compact, cdiv(n, 256), 256, 0, c->d_constrained, d_s, c->d_conlist, n);
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp); dfree(d_f); dfree(d_s);
}

// ---- interpolation ------------------------------------------------------------------------------------
struct InterpSpec { int n; int cond_kind[8]; uint32_t mask[8]; int child[8]; mdb_function fn[8]; };

// One thread per DOF of one child.  interpolate.hh:35-73 loops over the cells in host order and, inside a cell, over
// the (function, subproblem) pairs in argument order; the value that survives is the one written last, i.e. by the
// highest-index cell containing the DOF on which any listed subproblem applies, and there by the last such pair.
__global__ void k_interpolate(MeshDesc m, const uint8_t* __restrict__ sd, ChildDev ch, int64_t child_size, int child_index, InterpSpec spec,
                               double t, double* x)
{
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= child_size) return;
  int64_t r = ch.dof2lat[g];
  int p[3]; p[0] = (int)(r % ch.lat_n[0]); r /= ch.lat_n[0]; p[1] = (int)(r % ch.lat_n[1]); p[2] = (int)(r / ch.lat_n[1]);
  const int k = ch.k, st = ch.st;
  // candidate cells along each axis: c = floor(p/k) and, if p is on a cell boundary, c-1 (QkDG: the one cell that owns the point)
  int lo_c[3], hi_c[3];
  for (int d = 0; d < 3; ++d) {
    if (d >= m.dim) { lo_c[d] = hi_c[d] = 0; continue; }
    int c1 = p[d] / st;
    hi_c[d] = (c1 < m.n[d]) ? c1 : m.n[d] - 1;
    lo_c[d] = (!ch.dg && p[d] % k == 0 && c1 > 0) ? c1 - 1 : c1;
    if (lo_c[d] > hi_c[d]) lo_c[d] = hi_c[d];
  }
  for (int cz = hi_c[2]; cz >= lo_c[2]; --cz)
    for (int cy = hi_c[1]; cy >= lo_c[1]; --cy)
      for (int cx = hi_c[0]; cx >= lo_c[0]; --cx) {
        int ijk[3] = {cx, cy, cz};
        uint8_t s = sd[cell_index(m, ijk)];
        if (!child_on_cell(ch.subdomain, s)) continue;
        for (int q = spec.n - 1; q >= 0; --q) {
          if (spec.child[q] != child_index) continue;
          if (!cond_applies(spec.cond_kind[q], spec.mask[q], s)) continue;
          double xg[3];
          for (int d = 0; d < m.dim; ++d) {
            double lo = m.coord(d, ijk[d]), up = m.coord(d, ijk[d] + 1);
            double local = (double)(p[d] - st * ijk[d]) / k;
            xg[d] = lo + local * (up - lo);
          }
          x[ch.offset + g] = eval_function(spec.fn[q], m.dim, xg, t);
          return;
        }
      }
}

void interpolate(Ctx* c, const int* sps, const mdb_function* fns, int n, double t, double* d_x)
{
  MDB_REQUIRE(n <= 8, MDB_EINVAL, "at most 8 interpolation pairs");
  InterpSpec spec; spec.n = n;
  for (int q = 0; q < n; ++q) {
    MDB_REQUIRE(sps[q] >= 0 && sps[q] < (int)c->sps.size(), MDB_EINVAL, "bad subproblem index");
    const SubProblemDesc& sp = c->sps[sps[q]];
    spec.cond_kind[q] = sp.cond_kind; spec.mask[q] = sp.mask; spec.child[q] = sp.children[0]; spec.fn[q] = fns[q];
  }
  for (int i = 0; i < (int)c->children.size(); ++i) {
    const Child& ch = c->children[i];
    MDB_LAUNCH(c, k_interpolate, cdiv(ch.size, 256), 256, 0, c->mesh, c->d_cell_sd, child_dev(c, i), ch.size, i, spec, t, d_x);
  }
}

// ---- interface face lists ------------------------------------------------------------------------------
__global__ void k_face_flags(MeshDesc m, const uint8_t* __restrict__ sd, int lk, uint32_t lmask, int rk, uint32_t rmask, int32_t* flags)
{
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nf = 2 * m.dim;
  if (t >= m.ncells() * nf) return;
  int64_t c = t / nf; int f = (int)(t % nf);
  int ijk[3]; cell_ijk(m, c, ijk);
  int axis = f / 2, side = f % 2;
  int nijk[3] = {ijk[0], ijk[1], ijk[2]};
  nijk[axis] += side ? 1 : -1;
  int fl = 0;
  if (nijk[axis] >= 0 && nijk[axis] < m.n[axis])
    fl = cond_applies(lk, lmask, sd[c]) && cond_applies(rk, rmask, sd[cell_index(m, nijk)]);   // coupling.hh:78-83
  flags[t] = fl;
}

__global__ void k_face_compact(int nf, int64_t total, const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int32_t* cell, int8_t* face)
{
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= total || !flags[t]) return;
  cell[scan[t]] = (int32_t)(t / nf); face[scan[t]] = (int8_t)(t % nf);
}

// ---- which coefficients does jacobian(x, a) read? ------------------------------------------------------------------------
struct NeedSide { int k; int lat_n[3]; int64_t offset; const int32_t* lat2dof; int st; };
__global__ void k_mark_face_dofs(MeshDesc m, NeedSide L, NeedSide R, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces,
                                 uint8_t* flags)
{
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  int ijk[3]; cell_ijk(m, fcell[f], ijk);
  int nijk[3] = {ijk[0], ijk[1], ijk[2]};
  const int face = fface[f];
  nijk[face / 2] += (face % 2) ? 1 : -1;
  for (int side = 0; side < 2; ++side) {
    const NeedSide& S = side ? R : L;
    const int* c = side ? nijk : ijk;
    const int k = S.k, kz = (m.dim > 2) ? k : 0;
    for (int az = 0; az <= kz; ++az)
      for (int ay = 0; ay <= k; ++ay)
        for (int ax = 0; ax <= k; ++ax) {
          const int64_t p = (S.st * c[0] + ax) + (int64_t)S.lat_n[0] * ((S.st * c[1] + ay) + (int64_t)S.lat_n[1] * (S.st * c[2] + az));
          flags[S.offset + S.lat2dof[p]] = 1;
        }
  }
}

// the mortar DOFs of the coupling faces (coefficients x_c the enriched FD Jacobian reads)
__global__ void k_mark_face_mortar_dofs(MeshDesc m, NeedSide C, const int32_t* __restrict__ fcell, const int8_t* __restrict__ fface, int64_t nfaces,
                                        uint8_t* flags)
{
  int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= nfaces) return;
  int ijk[3]; cell_ijk(m, fcell[f], ijk);
  const int axis = fface[f] / 2, side = fface[f] % 2;
  for (int i = 0; i < (1 << (m.dim - 1)); ++i) {
    int v[3] = {ijk[0], ijk[1], ijk[2]}; int t = 0;
    for (int d = 0; d < m.dim; ++d) { if (d == axis) v[d] += side; else { v[d] += (i >> t) & 1; ++t; } }
    const int32_t g = C.lat2dof[v[0] + (int64_t)C.lat_n[0] * (v[1] + (int64_t)C.lat_n[1] * v[2])];
    if (g >= 0) flags[C.offset + g] = 1;
  }
}

void build_xneed(Ctx* c, GridOp& go)
{
  const int64_t n = c->ndof;
  uint8_t* d_flags = dalloc<uint8_t>(n);
  MDB_CUDA(cudaMemsetAsync(d_flags, 0, n, c->stream));
  for (size_t k = 0; k < go.cps.size(); ++k) {
    const CouplingDesc& cp = c->cps[go.cps[k]];
    const FaceList& fl = go.faces[k];
    if (fl.n == 0) continue;
    NeedSide S[2];
    const int sp[2] = {cp.local_sp, cp.remote_sp};
    for (int q = 0; q < 2; ++q) {
      const Child& ch = c->children[c->sps[sp[q]].children[0]];
      for (int d = 0; d < 3; ++d) S[q].lat_n[d] = ch.lat_n[d];
      S[q].k = ch.k; S[q].offset = ch.offset; S[q].lat2dof = ch.d_lat2dof; S[q].st = ch.stride();
    }
    MDB_LAUNCH(c, k_mark_face_dofs, cdiv(fl.n, 256), 256, 0, c->mesh, S[0], S[1], fl.d_cell, fl.d_face, fl.n, d_flags);
    for (int kc = 0; cp.enriched() && kc < cp.ncomp; ++kc) {
      const Child& ch = c->children[cp.coupling_child + kc];
      NeedSide C; for (int d = 0; d < 3; ++d) C.lat_n[d] = ch.lat_n[d];
      C.k = ch.k; C.offset = ch.offset; C.lat2dof = ch.d_lat2dof; C.st = ch.stride();
      MDB_LAUNCH(c, k_mark_face_mortar_dofs, cdiv(fl.n, 256), 256, 0, c->mesh, C, fl.d_cell, fl.d_face, fl.n, d_flags);
    }
  }
  int32_t* d_f = dalloc<int32_t>(n);
  int32_t* d_s = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_u8_to_i32, cdiv(n, 256), 256, 0, d_flags, d_f, n);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_f, d_s, (int)n, c->stream);
  void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_f, d_s, (int)n, c->stream)); c->launches++;
  int32_t ls = 0, lf = 0;
  MDB_CUDA(cudaMemcpyAsync(&ls, d_s + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&lf, d_f + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  go.n_xneed = (int64_t)ls + lf;
  dfree(go.d_xneed);
  go.d_xneed = dalloc<int32_t>(go.n_xneed);
  MDB_LAUNCH(c, k_compact, cdiv(n, 256), 256, 0, d_flags, d_s, go.d_xneed, n);
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp); dfree(d_f); dfree(d_s); dfree(d_flags);
}

void build_facelists(Ctx* c, GridOp& go)
{
  const MeshDesc& m = c->mesh;
  const int nf = 2 * m.dim;
  const int64_t total = m.ncells() * nf;
  MDB_REQUIRE(total < (int64_t)2147000000, MDB_EINVAL, "mesh too large for the face scan");
  go.faces.clear();
  if (go.cps.empty()) return;
  int32_t* d_f = dalloc<int32_t>(total);
  int32_t* d_s = dalloc<int32_t>(total);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_f, d_s, (int)total, c->stream);
  void* d_tmp = nullptr; MDB_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  for (int k : go.cps) {
    const CouplingDesc& cp = c->cps[k];
    const SubProblemDesc& l = c->sps[cp.local_sp]; const SubProblemDesc& r = c->sps[cp.remote_sp];
    MDB_LAUNCH(c, k_face_flags, cdiv(total, 256), 256, 0, m, c->d_cell_sd, l.cond_kind, l.mask, r.cond_kind, r.mask, d_f);
    MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_f, d_s, (int)total, c->stream)); c->launches++;
    int32_t ls = 0, lf = 0;
    MDB_CUDA(cudaMemcpyAsync(&ls, d_s + total - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaMemcpyAsync(&lf, d_f + total - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    FaceList fl; fl.n = (int64_t)ls + lf;
    fl.d_cell = dalloc<int32_t>(fl.n); fl.d_face = dalloc<int8_t>(fl.n);
    MDB_LAUNCH(c, k_face_compact, cdiv(total, 256), 256, 0, nf, total, d_f, d_s, fl.d_cell, fl.d_face);
    go.faces.push_back(fl);
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp); dfree(d_f); dfree(d_s);
}

} // namespace mdb
