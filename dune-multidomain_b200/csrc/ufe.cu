// ufe.cu — generic finite elements on an unstructured triangle mesh: the Stokes-Darcy path of the reference
// (test/stokesdarcy2.cc, SURVEY §8 rows a5 + a12) behind the same C ABI as the rest.
//
//   spaces      MDB_FE_P1 / MDB_FE_P2 (PkLocalFiniteElementMap, vertex + edge DOFs) and MDB_FE_OPBk (orthonormal basis, cell DOFs);
//               block numbering [UPSTREAM (ii),(iii)]: vertices < edges < cells, rank of the entity inside the child's subdomain
//   subproblems any number of children (subproblemlocalfunctionspace.hh:144-261): the local space of a subproblem on a cell is the
//               concatenation of its children's local spaces.  It is materialised ONCE per subproblem as a device table
//               lmap[cell][nloc] of container indices (the reference rebuilds and remaps it on every functor call);
//               every kernel below is written against that table and knows nothing about entities
//   pattern     keys (row << 32 | column) of the cell cliques (FullVolumePattern over all children), the DG skeleton links and
//               the coupling links (couplingutilities.hh:35-47) -> radix sort -> unique -> CSR: ascending columns = BCRS order
//   operators   MDB_OP_TAYLORHOOD_STOKES ([UPSTREAM] TaylorHoodNavierStokes, Stokes, full tensor), MDB_OP_DARCY_DG ([UPSTREAM]
//               ConvectionDiffusionDG(SIPG | NIPG, weights) with the layered tensor of DarcyParameters), MDB_OP_STOKESDARCY_COUPLING
//               (StokesDarcyCouplingOperator::alpha_coupling, test/stokesdarcycouplingoperator.hh:104-245, with the reference's
//               finite-difference Jacobian NumericalJacobianCoupling(epsilon), couplingutilities.hh:75-135)
//   work items  Taylor-Hood: a thread per (cell, local row) adds its row of the element matrix by atomics (rows are shared between
//               cells); DG: a thread per (cell, local row) OWNS its matrix row (volume + the three faces seen from this cell — the
//               interior-penalty form is symmetric in the two sides, so no face is visited from "the other" side): no atomics;
//               coupling: a thread per (interface face, perturbed coefficient)
// Basis functions are tabulated once on the host at the quadrature points of the reference element (same operation order as
// the test oracle's restatement) and read from global memory.  Compiled with --fmad=false (reference operation order for the FD Jacobian).
#include "common.cuh"
#include "opb_tables.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>

namespace mdb {

namespace {

constexpr int UFE_MAXN = 10;      // local basis size of one child on a cell
constexpr int UFE_MAXQ = 16;      // volume points: order 6 -> 4 x 4 collapsed Gauss points
constexpr int UFE_MAXQF = 5;
constexpr int UFE_MAXLOC = 16;    // local size of a subproblem (Taylor-Hood: 6 + 6 + 3)
constexpr uint64_t UFE_NOKEY = ~0ull;

struct FeVol { int n, nq; double w[UFE_MAXQ]; double x[UFE_MAXQ][2]; double val[UFE_MAXQ][UFE_MAXN]; double grad[UFE_MAXQ][UFE_MAXN][2]; };
// face tables for the ordered vertex pair (a, b) of the reference triangle, index 3 a + b: point = corner a + t (corner b - corner a)
struct FeFace { int n, nq; double w[UFE_MAXQF], t[UFE_MAXQF]; double val[9][UFE_MAXQF][UFE_MAXN]; double grad[9][UFE_MAXQF][UFE_MAXN][2]; };

struct UfeTables {
  std::map<std::pair<int, int>, FeVol*> vol;      // (fe, order) -> device table
  std::map<std::pair<int, int>, FeFace*> face;
  ~UfeTables() { for (auto& p : vol) cudaFree(p.second); for (auto& p : face) cudaFree(p.second); }
};

int fe_size(int fe) { return fe == MDB_FE_P1 ? 3 : fe == MDB_FE_P2 ? 6 : fe == MDB_FE_OPB1 ? 3 : fe == MDB_FE_OPB2 ? 6 : 10; }
int fe_order(int fe) { return fe == MDB_FE_P1 ? 1 : fe == MDB_FE_P2 ? 2 : fe - 30; }

// ---- local bases on the host (same operation order as the test oracle: pk_eval, opb_eval) --------------------------------
// [UPSTREAM] dune-localfunctions Pk2DLocalBasis: Lagrange basis on the lattice (i/k, j/k), i fastest
void pk_eval_host(int k, double x, double y, double* val, double (*grad)[2])
{
  double pos[4];
  for (int a = 0; a <= k; ++a) pos[a] = a / (double)k;
  int n = 0;
  for (int j = 0; j <= k; ++j)
    for (int i = 0; i <= k - j; ++i) {
      double f0[6] = {0}, f1[6] = {0}, f2[6] = {0}; int nf = 0;
      for (int a = 0; a < i; ++a) { f0[nf] = (x - pos[a]) / (pos[i] - pos[a]); f1[nf] = 1.0 / (pos[i] - pos[a]); f2[nf] = 0.0; ++nf; }
      for (int b = 0; b < j; ++b) { f0[nf] = (y - pos[b]) / (pos[j] - pos[b]); f1[nf] = 0.0; f2[nf] = 1.0 / (pos[j] - pos[b]); ++nf; }
      for (int c = i + j + 1; c <= k; ++c) { const double den = pos[c] - pos[i] - pos[j]; f0[nf] = (pos[c] - x - y) / den; f1[nf] = -1.0 / den; f2[nf] = -1.0 / den; ++nf; }
      double v = 1.0;
      for (int l = 0; l < nf; ++l) v *= f0[l];
      double gx = 0.0, gy = 0.0;
      for (int m = 0; m < nf; ++m) {
        double px = 1.0, py = 1.0;
        for (int l = 0; l < nf; ++l) { if (l == m) { px *= f1[l]; py *= f2[l]; } else { px *= f0[l]; py *= f0[l]; } }
        gx += px; gy += py;
      }
      val[n] = v; grad[n][0] = gx; grad[n][1] = gy; ++n;
    }
}
void opb_eval_host(int k, double x, double y, double* val, double (*grad)[2])
{
  const int n = (k + 1) * (k + 2) / 2;
  const int (*ex)[2] = k == 1 ? opb1_exp : k == 2 ? opb2_exp : opb3_exp;
  auto L = [&](int i, int j) { return k == 1 ? opb1_L[i][j] : k == 2 ? opb2_L[i][j] : opb3_L[i][j]; };
  double px[4] = {1.0, 0, 0, 0}, py[4] = {1.0, 0, 0, 0};
  for (int a = 1; a <= k; ++a) { px[a] = px[a - 1] * x; py[a] = py[a - 1] * y; }
  double m[10], mx[10], my[10];
  for (int j = 0; j < n; ++j) {
    const int a = ex[j][0], b = ex[j][1];
    m[j] = px[a] * py[b];
    mx[j] = a > 0 ? ((double)a * px[a - 1]) * py[b] : 0.0;
    my[j] = b > 0 ? ((double)b * px[a]) * py[b - 1] : 0.0;
  }
  for (int i = 0; i < n; ++i) {
    double v = 0.0, gx = 0.0, gy = 0.0;
    for (int j = 0; j <= i; ++j) { v += L(i, j) * m[j]; gx += L(i, j) * mx[j]; gy += L(i, j) * my[j]; }
    val[i] = v; grad[i][0] = gx; grad[i][1] = gy;
  }
}
void fe_eval_host(int fe, double x, double y, double* val, double (*grad)[2])
{
  if (fe == MDB_FE_P1) pk_eval_host(1, x, y, val, grad);
  else if (fe == MDB_FE_P2) pk_eval_host(2, x, y, val, grad);
  else opb_eval_host(fe - 30, x, y, val, grad);
}

const double h_ref[3][2] = {{0.0, 0.0}, {1.0, 0.0}, {0.0, 1.0}};

UfeTables* tables(Ctx* c)
{
  if (!c->um->ufe_tables) c->um->ufe_tables = new UfeTables();
  return (UfeTables*)c->um->ufe_tables;
}
// volume table: collapsed Gauss product rule exact to degree `order` (ceil((order + 2) / 2) points per direction)
const FeVol* vol_table(Ctx* c, int fe, int order)
{
  UfeTables* T = tables(c);
  auto it = T->vol.find({fe, order});
  if (it != T->vol.end()) return it->second;
  const Rule1D g = gauss1d(order + 1);
  MDB_REQUIRE(g.m * g.m <= UFE_MAXQ, MDB_EINVAL, "generic elements: triangle rules up to order 6");
  FeVol h; std::memset(&h, 0, sizeof(h));
  h.n = fe_size(fe); h.nq = g.m * g.m;
  int q = 0;
  for (int a = 0; a < g.m; ++a)
    for (int b = 0; b < g.m; ++b, ++q) {
      const double s = g.x[a], t = g.x[b];
      h.x[q][0] = s; h.x[q][1] = t * (1.0 - s);
      h.w[q] = g.w[a] * g.w[b] * (1.0 - s);
      fe_eval_host(fe, h.x[q][0], h.x[q][1], h.val[q], h.grad[q]);
    }
  FeVol* d = nullptr; MDB_CUDA(cudaMalloc(&d, sizeof(FeVol)));
  MDB_CUDA(cudaMemcpy(d, &h, sizeof(FeVol), cudaMemcpyHostToDevice));
  T->vol[{fe, order}] = d;
  return d;
}
const FeFace* face_table(Ctx* c, int fe, int order)
{
  UfeTables* T = tables(c);
  auto it = T->face.find({fe, order});
  if (it != T->face.end()) return it->second;
  const Rule1D g = gauss1d(order);
  MDB_REQUIRE(g.m <= UFE_MAXQF, MDB_EINVAL, "generic elements: face rules up to order 9");
  FeFace h; std::memset(&h, 0, sizeof(h));
  h.n = fe_size(fe); h.nq = g.m;
  for (int q = 0; q < g.m; ++q) { h.w[q] = g.w[q]; h.t[q] = g.x[q]; }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      if (a == b) continue;
      for (int q = 0; q < g.m; ++q) {
        const double x = h_ref[a][0] + g.x[q] * (h_ref[b][0] - h_ref[a][0]), y = h_ref[a][1] + g.x[q] * (h_ref[b][1] - h_ref[a][1]);
        fe_eval_host(fe, x, y, h.val[3 * a + b][q], h.grad[3 * a + b][q]);
      }
    }
  FeFace* d = nullptr; MDB_CUDA(cudaMalloc(&d, sizeof(FeFace)));
  MDB_CUDA(cudaMemcpy(d, &h, sizeof(FeFace), cudaMemcpyHostToDevice));
  T->face[{fe, order}] = d;
  return d;
}

// ---- device views ------------------------------------------------------------------------------------------------------------------
struct UM { int64_t nv, ne; const double* xy; const int32_t *conn, *nbr, *cell_edge, *group; const uint8_t* sd; };
struct UChild { int fe, nloc, subdomain; int64_t offset, nvs; const int32_t *v2d, *e2d, *c2d; };
struct USp { int cond_kind; uint32_t mask; int nkids, nloc; UChild kid[4]; };

UM um_view(const Ctx* c)
{
  const UMesh* u = c->um;
  UM m; m.nv = u->nv; m.ne = u->ne; m.xy = u->d_xy; m.conn = u->d_conn; m.nbr = u->d_nbr; m.cell_edge = u->d_cell_edge; m.group = u->d_group; m.sd = c->d_cell_sd;
  return m;
}
UChild child_view(const Child& ch)
{
  UChild k; k.fe = ch.fe_kind; k.nloc = ch.um_nloc; k.subdomain = ch.subdomain; k.offset = ch.offset; k.nvs = ch.nvs;
  k.v2d = ch.d_lat2dof; k.e2d = ch.d_edge2dof; k.c2d = ch.d_cell2dof;
  return k;
}
USp sp_view(const Ctx* c, const SubProblemDesc& d)
{
  USp s; s.cond_kind = d.cond_kind; s.mask = d.mask; s.nkids = (int)d.children.size(); s.nloc = 0;
  for (int i = 0; i < s.nkids; ++i) { s.kid[i] = child_view(c->children[d.children[i]]); s.nloc += s.kid[i].nloc; }
  return s;
}

// container index of local function i of child k on cell e (multidomainlocalfunctionspace.hh:348-358)
__device__ inline int32_t uchild_dof(const UM& m, const UChild& k, int e, int i)
{
  if (k.fe == MDB_FE_P1) return (int32_t)(k.offset + k.v2d[m.conn[3 * e + i]]);
  if (k.fe == MDB_FE_P2) {
    // Pk2D lattice order: vertex 0, edge {0,1}, vertex 1, edge {0,2}, edge {1,2}, vertex 2
    switch (i) {
    case 0: return (int32_t)(k.offset + k.v2d[m.conn[3 * e]]);
    case 2: return (int32_t)(k.offset + k.v2d[m.conn[3 * e + 1]]);
    case 5: return (int32_t)(k.offset + k.v2d[m.conn[3 * e + 2]]);
    case 1: return (int32_t)(k.offset + k.nvs + k.e2d[m.cell_edge[3 * e]]);
    case 3: return (int32_t)(k.offset + k.nvs + k.e2d[m.cell_edge[3 * e + 1]]);
    default: return (int32_t)(k.offset + k.nvs + k.e2d[m.cell_edge[3 * e + 2]]);
    }
  }
  return (int32_t)(k.offset + (int64_t)k.c2d[e] * k.nloc + i);
}

// ---- DOF map ---------------------------------------------------------------------------------------------------------------------------
__global__ void k_ufe_mark(UM m, int subdomain, int32_t* vflag, int32_t* eflag, int32_t* cflag)
{
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= m.ne) return;
  const int in = subdomain < 0 || ((m.sd[e] >> subdomain) & 1);
  cflag[e] = in;
  if (!in) return;
  for (int i = 0; i < 3; ++i) { vflag[m.conn[3 * e + i]] = 1; eflag[m.cell_edge[3 * e + i]] = 1; }
}
__global__ void k_ufe_assign(int64_t n, const int32_t* flag, const int32_t* scan, int32_t* map)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) map[i] = flag[i] ? scan[i] : -1;
}
__global__ void k_ufe_lmap(UM m, USp s, int32_t* lmap)
{
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= m.ne) return;
  const bool on = cond_applies(s.cond_kind, s.mask, m.sd[e]);
  int o = 0;
  for (int k = 0; k < s.nkids; ++k)
    for (int i = 0; i < s.kid[k].nloc; ++i, ++o) lmap[e * (int64_t)s.nloc + o] = on ? uchild_dof(m, s.kid[k], (int)e, i) : -1;
}

int64_t scan_i32(Ctx* c, const int32_t* d_in, int32_t* d_out, int64_t n)
{
  if (n == 0) return 0;
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_in, d_out, (int)n, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_in, d_out, (int)n, c->stream)); c->launches++;
  int32_t last_s = 0, last_f = 0;
  MDB_CUDA(cudaMemcpyAsync(&last_s, d_out + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaMemcpyAsync(&last_f, d_in + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_t);
  return (int64_t)last_s + last_f;
}

// ---- constraints -------------------------------------------------------------------------------------------------------------------------
struct UCon { int n; const int32_t* lmap[8]; int nloc[8]; int nkids[8]; int kid_fe[8][4]; mdb_bctype bc[8]; };
__device__ __constant__ int c_ufe_face_v[3][2] = {{0, 1}, {0, 2}, {1, 2}};
__device__ __constant__ int c_ufe_p2_face[3][3] = {{0, 1, 2}, {0, 3, 5}, {2, 4, 5}};     // P2 functions attached to the sub-entities of face f

// constraintsfunctors.hh:75-103 with [UPSTREAM] ConformingDirichletConstraints / StokesVelocityDirichletConstraints: on a face whose
// outside cell is not in the subproblem the predicate is asked at the face centre; a Dirichlet face constrains the coefficients
// attached to its sub-entities (the listed children only: the Taylor-Hood pressure is never constrained)
__global__ void k_ufe_constraints(UM m, UCon B, uint8_t* constrained)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m.ne) return;
  const int e = (int)(i / 3), f = (int)(i % 3);
  const int nb = m.nbr[i];
  const int va = m.conn[3 * e + c_ufe_face_v[f][0]], vb = m.conn[3 * e + c_ufe_face_v[f][1]];
  double xg[2];
  for (int d = 0; d < 2; ++d) { const double pa = m.xy[2 * va + d], pb = m.xy[2 * vb + d]; xg[d] = pa + 0.5 * (pb - pa); }
  for (int q = 0; q < B.n; ++q) {
    const int32_t* row = B.lmap[q] + (int64_t)e * B.nloc[q];
    if (row[0] < 0) continue;
    if (nb >= 0 && B.lmap[q][(int64_t)nb * B.nloc[q]] >= 0) continue;
    if (eval_bctype(B.bc[q], 2, xg, nb < 0) != 0) continue;
    int o = 0;
    for (int k = 0; k < B.nkids[q]; ++k) {
      const int fe = B.kid_fe[q][k];
      if (fe == MDB_FE_P2) { for (int a = 0; a < 3; ++a) constrained[row[o + c_ufe_p2_face[f][a]]] = 1; o += 6; }
      else if (fe == MDB_FE_P1) { constrained[row[o + c_ufe_face_v[f][0]]] = 1; constrained[row[o + c_ufe_face_v[f][1]]] = 1; o += 3; }
      else o += (fe - 30 + 1) * (fe - 30 + 2) / 2;                  // cell-attached coefficients: nothing on a face
    }
  }
}

// ---- coupling faces ------------------------------------------------------------------------------------------------------------------------
__global__ void k_ufe_face_flags(UM m, const int32_t* lmapL, int nL, const int32_t* lmapR, int nR, int32_t* flag)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m.ne) return;
  const int nb = m.nbr[i];
  flag[i] = (nb >= 0 && lmapL[(i / 3) * nL] >= 0 && lmapR[(int64_t)nb * nR] >= 0) ? 1 : 0;     // coupling.hh:78-83: local on the inside, remote on the outside
}
__global__ void k_ufe_face_compact(int64_t n, const int32_t* flag, const int32_t* scan, int32_t* cell, int8_t* face)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  cell[scan[i]] = (int32_t)(i / 3); face[scan[i]] = (int8_t)(i % 3);
}

// ---- pattern ---------------------------------------------------------------------------------------------------------------------------------
__global__ void k_ufe_keys_volume(UM m, const int32_t* __restrict__ lmap, int nloc, int skel, uint64_t* keys)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= m.ne * nloc) return;
  const int64_t e = idx / nloc;
  const int32_t row = lmap[idx];
  uint64_t* out = keys + idx * (int64_t)nloc * (skel ? 4 : 1);
  for (int j = 0; j < nloc; ++j) out[j] = row >= 0 ? (((uint64_t)row << 32) | (uint32_t)lmap[e * nloc + j]) : UFE_NOKEY;
  if (!skel) return;
  for (int f = 0; f < 3; ++f) {                      // doPatternSkeleton: FullSkeletonPattern on the faces inside the subproblem
    const int nb = m.nbr[3 * e + f];
    const bool ok = row >= 0 && nb >= 0 && lmap[(int64_t)nb * nloc] >= 0;
    for (int j = 0; j < nloc; ++j) out[(1 + f) * nloc + j] = ok ? (((uint64_t)row << 32) | (uint32_t)lmap[(int64_t)nb * nloc + j]) : UFE_NOKEY;
  }
}
// pattern_coupling: lfsv_s x lfsu_n and lfsv_n x lfsu_s of every face the coupling fires on (couplingutilities.hh:35-47)
__global__ void k_ufe_keys_coupling(UM m, const int32_t* __restrict__ lmapL, int nL, const int32_t* __restrict__ lmapR, int nR,
                                    const int32_t* fcell, const int8_t* fface, int64_t nf, uint64_t* keys)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nf * (nL + nR)) return;
  const int64_t fi = idx / (nL + nR); const int i = (int)(idx % (nL + nR));
  const int e = fcell[fi], nb = m.nbr[3 * e + fface[fi]];
  uint64_t* out = keys + fi * 2 * (int64_t)nL * nR;
  if (i < nL) { const uint64_t row = (uint32_t)lmapL[(int64_t)e * nL + i]; for (int j = 0; j < nR; ++j) out[i * nR + j] = (row << 32) | (uint32_t)lmapR[(int64_t)nb * nR + j]; }
  else { const uint64_t row = (uint32_t)lmapR[(int64_t)nb * nR + (i - nL)]; for (int j = 0; j < nL; ++j) out[nL * nR + (i - nL) * nL + j] = (row << 32) | (uint32_t)lmapL[(int64_t)e * nL + j]; }
}
__global__ void k_ufe_csr(const uint64_t* __restrict__ keys, int64_t nnz, int64_t ndof, int64_t* rowptr, int32_t* colidx)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nnz) return;
  const int64_t row = (int64_t)(keys[i] >> 32);
  colidx[i] = (int32_t)(keys[i] & 0xffffffffu);
  const int64_t prev = i > 0 ? (int64_t)(keys[i - 1] >> 32) : -1;
  for (int64_t r = prev + 1; r <= row; ++r) rowptr[r] = i;
  if (i == nnz - 1) for (int64_t r = row + 1; r <= ndof; ++r) rowptr[r] = nnz;
}
__device__ inline int ufe_find(const int32_t* cols, int len, int32_t col)
{
  int lo = 0, hi = len - 1;
  while (lo <= hi) { const int mid = (lo + hi) >> 1; const int32_t v = cols[mid]; if (v == col) return mid; if (v < col) lo = mid + 1; else hi = mid - 1; }
  return -1;
}
__global__ void k_ufe_diag(int64_t ndof, const int64_t* rowptr, const int32_t* colidx, int32_t* diag, int32_t* len)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ndof) return;
  const int64_t p0 = rowptr[r];
  len[r] = (int32_t)(rowptr[r + 1] - p0);
  diag[r] = ufe_find(colidx + p0, len[r], (int32_t)r);
}

// ---- geometry ----------------------------------------------------------------------------------------------------------------------------------
struct UTri { double p[3][2]; double t[2][2]; double ie; };      // x = p0 + J xi; t = J^{-T}; ie = |det J|
__device__ inline UTri utri(const UM& m, int e)
{
  UTri g;
  for (int i = 0; i < 3; ++i) { const int v = m.conn[3 * e + i]; g.p[i][0] = m.xy[2 * v]; g.p[i][1] = m.xy[2 * v + 1]; }
  const double ax = g.p[1][0] - g.p[0][0], ay = g.p[1][1] - g.p[0][1];
  const double bx = g.p[2][0] - g.p[0][0], by = g.p[2][1] - g.p[0][1];
  const double det = ax * by - ay * bx;
  const double inv = 1.0 / det;
  g.t[0][0] = by * inv; g.t[0][1] = -(ay * inv);
  g.t[1][0] = -(bx * inv); g.t[1][1] = ax * inv;
  g.ie = fabs(det);
  return g;
}
struct UFaceGeo { int la, lb; double pa[2], d[2], ie, n[2]; };   // face f of cell e: corner a, edge vector, length, unit outer normal
__device__ inline UFaceGeo uface(const UTri& g, int f)
{
  UFaceGeo F; F.la = c_ufe_face_v[f][0]; F.lb = c_ufe_face_v[f][1];
  F.pa[0] = g.p[F.la][0]; F.pa[1] = g.p[F.la][1];
  F.d[0] = g.p[F.lb][0] - F.pa[0]; F.d[1] = g.p[F.lb][1] - F.pa[1];
  F.ie = sqrt(F.d[0] * F.d[0] + F.d[1] * F.d[1]);
  double nx = F.d[1] / F.ie, ny = -F.d[0] / F.ie;
  const int o = 2 - f;                                           // the vertex opposite to face f
  if (nx * (g.p[o][0] - F.pa[0]) + ny * (g.p[o][1] - F.pa[1]) > 0.0) { nx = -nx; ny = -ny; }
  F.n[0] = nx; F.n[1] = ny;
  return F;
}
// index 3 a + b of the ordered pair of the NEIGHBOUR's local vertices that are the face's corners a, b (geometryInOutside)
__device__ inline int uface_outside_pair(const UM& m, int e, int f, int nb)
{
  const int va = m.conn[3 * e + c_ufe_face_v[f][0]], vb = m.conn[3 * e + c_ufe_face_v[f][1]];
  int ia = 0, ib = 0;
  for (int k = 0; k < 3; ++k) { const int v = m.conn[3 * nb + k]; if (v == va) ia = k; if (v == vb) ib = k; }
  return 3 * ia + ib;
}
__device__ inline void ugrad(const UTri& g, const double js[2], double out[2])
{
  out[0] = g.t[0][0] * js[0] + g.t[0][1] * js[1];
  out[1] = g.t[1][0] * js[0] + g.t[1][1] * js[1];
}

// ---- [UPSTREAM] TaylorHoodNavierStokes (Stokes, optional full tensor): local space [v_0 (6) | v_1 (6) | p (3)] -----------------------
struct UTh { mdb_taylorhood_params P; const FeVol* v2; const FeVol* p1; const FeFace* v2f; };

// row i of the element matrix (the operator is linear: jacobian_volume does not read x)
__global__ void k_ufe_th_jacobian(UM m, UTh T, const int32_t* __restrict__ lmap, double weight, const int64_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ colidx, double* val)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= m.ne * 15) return;
  const int e = (int)(idx / 15), i = (int)(idx % 15);
  const int32_t* lm = lmap + (int64_t)e * 15;
  if (lm[0] < 0) return;
  const UTri g = utri(m, e);
  double K[15];
  for (int j = 0; j < 15; ++j) K[j] = 0.0;
  const double mu = T.P.mu;
  for (int q = 0; q < T.v2->nq; ++q) {
    const double factor = T.v2->w[q] * g.ie;
    double gr[6][2];
    for (int b = 0; b < 6; ++b) ugrad(g, T.v2->grad[q][b], gr[b]);
    if (i < 12) {
      const int d = i / 6, a = i % 6;
      for (int b = 0; b < 6; ++b) {
        K[6 * d + b] += mu * (gr[b][0] * gr[a][0] + gr[b][1] * gr[a][1]) * factor;
        if (T.P.full_tensor) for (int dd = 0; dd < 2; ++dd) K[6 * dd + b] += mu * gr[b][d] * gr[a][dd] * factor;
      }
      for (int b = 0; b < 3; ++b) K[12 + b] += -(T.p1->val[q][b] * gr[a][d]) * factor;
    } else {
      const double psi = T.p1->val[q][i - 12];
      for (int dd = 0; dd < 2; ++dd) for (int b = 0; b < 6; ++b) K[6 * dd + b] += -(gr[b][dd] * psi) * factor;
    }
  }
  const int64_t p0 = rowptr[lm[i]]; const int len = (int)(rowptr[lm[i] + 1] - p0);
  for (int j = 0; j < 15; ++j) {
    if (i >= 12 && j >= 12) continue;                           // the pressure block is stored and stays zero
    const int pos = ufe_find(colidx + p0, len, lm[j]);
    if (pos >= 0) atomicAdd(&val[p0 + pos], weight * K[j]);
  }
}
// alpha_volume + lambda_volume + lambda_boundary, row i
__global__ void k_ufe_th_residual(UM m, UTh T, const int32_t* __restrict__ lmap, double weight, const double* __restrict__ x, double* r)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= m.ne * 15) return;
  const int e = (int)(idx / 15), i = (int)(idx % 15);
  const int32_t* lm = lmap + (int64_t)e * 15;
  if (lm[0] < 0) return;
  const UTri g = utri(m, e);
  double xl[15];
  for (int j = 0; j < 15; ++j) xl[j] = x[lm[j]];
  const double mu = T.P.mu;
  double acc = 0.0;
  for (int q = 0; q < T.v2->nq; ++q) {
    const double factor = T.v2->w[q] * g.ie;
    double gr[6][2];
    for (int b = 0; b < 6; ++b) ugrad(g, T.v2->grad[q][b], gr[b]);
    double jacu[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int d = 0; d < 2; ++d) for (int b = 0; b < 6; ++b) { jacu[d][0] += xl[6 * d + b] * gr[b][0]; jacu[d][1] += xl[6 * d + b] * gr[b][1]; }
    if (i < 12) {
      const int d = i / 6, a = i % 6;
      double pres = 0.0;
      for (int b = 0; b < 3; ++b) pres += xl[12 + b] * T.p1->val[q][b];
      acc += (mu * (gr[a][0] * jacu[d][0] + gr[a][1] * jacu[d][1]) - pres * gr[a][d]) * factor;
      if (T.P.full_tensor) for (int dd = 0; dd < 2; ++dd) acc += mu * jacu[dd][d] * gr[a][dd] * factor;
      acc += -T.P.f[d] * T.v2->val[q][a] * factor;
    } else {
      const double divu = jacu[0][0] + jacu[1][1];
      acc += -divu * T.p1->val[q][i - 12] * factor;
    }
  }
  if (i < 12) {
    const int d = i / 6, a = i % 6;
    for (int f = 0; f < 3; ++f) {
      if (m.nbr[3 * e + f] >= 0) continue;                       // interior faces (subdomain interfaces included): DoNothing
      const UFaceGeo F = uface(g, f);
      for (int q = 0; q < T.v2f->nq; ++q) {
        const double tq = T.v2f->t[q];
        double xg[2] = {F.pa[0] + tq * F.d[0], F.pa[1] + tq * F.d[1]};
        if (eval_bctype(T.P.bc, 2, xg, true) != 1) continue;
        const double jv = eval_function(T.P.j, 2, xg, 0.0);
        acc += (jv * F.n[d]) * T.v2f->val[3 * F.la + F.lb][q][a] * (T.v2f->w[q] * F.ie);
      }
    }
  }
  atomicAdd(&r[lm[i]], weight * acc);
}

// ---- [UPSTREAM] ConvectionDiffusionDG with DarcyParameters on an orthonormal Pk space -----------------------------------------------------
struct UDg { mdb_darcy_params P; const FeVol* vol; const FeFace* face; int nk, k; };
__device__ inline double udg_perm(const UDg& D, const UM& m, int e) { return m.group[e] == D.P.high_group ? D.P.kabs : D.P.kabs_low; }

struct UDgFace { double omega_s, omega_n, penalty, As, An; int bc; int pair_s, pair_n; };
// face data seen from cell e (inside); nb < 0: boundary face (type at the face centre)
__device__ inline UDgFace udg_face(const UDg& D, const UM& m, int e, int f, int nb, const UTri& gs, const UFaceGeo& F)
{
  UDgFace R;
  R.As = udg_perm(D, m, e); R.An = 0.0;
  R.pair_s = 3 * F.la + F.lb; R.pair_n = 0;
  const double delta_s = (R.As * F.n[0]) * F.n[0] + (R.As * F.n[1]) * F.n[1];
  double harmonic, vol;
  if (nb < 0) {
    double xc[2] = {F.pa[0] + 0.5 * F.d[0], F.pa[1] + 0.5 * F.d[1]};
    R.bc = eval_bctype(D.P.bc, 2, xc, true);
    R.omega_s = 1.0; R.omega_n = 0.0; harmonic = delta_s; vol = 0.5 * gs.ie;
    if (!D.P.weights) harmonic = 1.0;
  } else {
    R.bc = -1;
    R.An = udg_perm(D, m, nb);
    const UTri gn = utri(m, nb);
    const double delta_n = (R.An * F.n[0]) * F.n[0] + (R.An * F.n[1]) * F.n[1];
    if (D.P.weights) {
      R.omega_s = delta_n / (delta_s + delta_n + 1e-20); R.omega_n = delta_s / (delta_s + delta_n + 1e-20);
      harmonic = 2.0 * delta_s * delta_n / (delta_s + delta_n + 1e-20);
    } else { R.omega_s = R.omega_n = 0.5; harmonic = 1.0; }
    vol = fmin(0.5 * gs.ie, 0.5 * gn.ie);
    R.pair_n = uface_outside_pair(m, e, f, nb);
  }
  const double h_F = vol / F.ie;
  R.penalty = (D.P.alpha / h_F) * harmonic * D.k * (D.k + 2 - 1);
  return R;
}

// The thread of (cell e, local row i) owns matrix row (e, i): volume block + the faces of e seen from e
__global__ void k_ufe_dg_jacobian(UM m, UDg D, const int32_t* __restrict__ lmap, double weight, const int64_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ colidx, double* val)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nk = D.nk;
  if (idx >= m.ne * nk) return;
  const int e = (int)(idx / nk), i = (int)(idx % nk);
  const int32_t* lm = lmap + (int64_t)e * nk;
  if (lm[0] < 0) return;
  const UTri g = utri(m, e);
  const double theta = D.P.method == MDB_DG_METHOD_NIPG ? 1.0 : -1.0;
  const double As = udg_perm(D, m, e);
  double own[UFE_MAXN];
  for (int j = 0; j < nk; ++j) own[j] = 0.0;
  for (int q = 0; q < D.vol->nq; ++q) {
    const double factor = D.vol->w[q] * g.ie;
    double gi[2]; ugrad(g, D.vol->grad[q][i], gi);
    for (int j = 0; j < nk; ++j) { double gj[2]; ugrad(g, D.vol->grad[q][j], gj); own[j] += ((As * gj[0]) * gi[0] + (As * gj[1]) * gi[1]) * factor; }
  }
  const int64_t p0 = rowptr[lm[i]]; const int len = (int)(rowptr[lm[i] + 1] - p0);
  for (int f = 0; f < 3; ++f) {
    const int nb = m.nbr[3 * e + f];
    if (nb >= 0 && lmap[(int64_t)nb * nk] < 0) continue;          // subdomain interface: boundary method with BCType None
    const UFaceGeo F = uface(g, f);
    const UDgFace R = udg_face(D, m, e, f, nb, g, F);
    if (nb < 0 && R.bc != 0) continue;                             // Neumann / outflow (b = 0): no Jacobian term
    const double Ans[2] = {R.As * F.n[0], R.As * F.n[1]};
    double nbk[UFE_MAXN];
    for (int j = 0; j < nk; ++j) nbk[j] = 0.0;
    UTri gn; double Ann[2] = {0.0, 0.0};
    if (nb >= 0) { gn = utri(m, nb); Ann[0] = R.An * F.n[0]; Ann[1] = R.An * F.n[1]; }
    for (int q = 0; q < D.face->nq; ++q) {
      const double factor = D.face->w[q] * F.ie;
      const double phi_i = D.face->val[R.pair_s][q][i];
      double gi[2]; ugrad(g, D.face->grad[R.pair_s][q][i], gi);
      const double gi_An = gi[0] * Ans[0] + gi[1] * Ans[1];
      for (int j = 0; j < nk; ++j) {
        double gj[2]; ugrad(g, D.face->grad[R.pair_s][q][j], gj);
        const double phi_j = D.face->val[R.pair_s][q][j];
        own[j] += (-(R.omega_s * (Ans[0] * gj[0] + Ans[1] * gj[1])) * factor) * phi_i + (phi_j * factor) * theta * R.omega_s * gi_An + (R.penalty * phi_j * factor) * phi_i;
      }
      if (nb >= 0)
        for (int j = 0; j < nk; ++j) {
          double gj[2]; ugrad(gn, D.face->grad[R.pair_n][q][j], gj);
          const double phi_j = D.face->val[R.pair_n][q][j];
          nbk[j] += (-(R.omega_n * (Ann[0] * gj[0] + Ann[1] * gj[1])) * factor) * phi_i - (phi_j * factor) * theta * R.omega_s * gi_An - (R.penalty * phi_j * factor) * phi_i;
        }
    }
    if (nb >= 0) {
      const int pos = ufe_find(colidx + p0, len, lmap[(int64_t)nb * nk]);       // the cell's coefficients are consecutive columns
      if (pos >= 0) for (int j = 0; j < nk; ++j) val[p0 + pos + j] += weight * nbk[j];
    }
  }
  const int pos = ufe_find(colidx + p0, len, lm[0]);
  if (pos >= 0) for (int j = 0; j < nk; ++j) val[p0 + pos + j] += weight * own[j];
}
__global__ void k_ufe_dg_residual(UM m, UDg D, const int32_t* __restrict__ lmap, double weight, const double* __restrict__ x, double* r)
{
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int nk = D.nk;
  if (idx >= m.ne * nk) return;
  const int e = (int)(idx / nk), i = (int)(idx % nk);
  const int32_t* lm = lmap + (int64_t)e * nk;
  if (lm[0] < 0) return;
  const UTri g = utri(m, e);
  const double theta = D.P.method == MDB_DG_METHOD_NIPG ? 1.0 : -1.0;
  const double As = udg_perm(D, m, e);
  double xs[UFE_MAXN];
  for (int j = 0; j < nk; ++j) xs[j] = x[lm[j]];
  double acc = 0.0;
  for (int q = 0; q < D.vol->nq; ++q) {
    const double factor = D.vol->w[q] * g.ie;
    double gi[2]; ugrad(g, D.vol->grad[q][i], gi);
    double gu[2] = {0.0, 0.0};
    for (int j = 0; j < nk; ++j) { double gj[2]; ugrad(g, D.vol->grad[q][j], gj); gu[0] += xs[j] * gj[0]; gu[1] += xs[j] * gj[1]; }
    acc += (gi[0] * (As * gu[0]) + gi[1] * (As * gu[1])) * factor;
  }
  for (int f = 0; f < 3; ++f) {
    const int nb = m.nbr[3 * e + f];
    if (nb >= 0 && lmap[(int64_t)nb * nk] < 0) continue;
    const UFaceGeo F = uface(g, f);
    const UDgFace R = udg_face(D, m, e, f, nb, g, F);
    const double Ans[2] = {R.As * F.n[0], R.As * F.n[1]};
    UTri gn; double Ann[2] = {0.0, 0.0}; double xn[UFE_MAXN];
    if (nb >= 0) { gn = utri(m, nb); Ann[0] = R.An * F.n[0]; Ann[1] = R.An * F.n[1]; for (int j = 0; j < nk; ++j) xn[j] = x[lmap[(int64_t)nb * nk + j]]; }
    for (int q = 0; q < D.face->nq; ++q) {
      const double factor = D.face->w[q] * F.ie;
      const double tq = D.face->t[q];
      const double phi_i = D.face->val[R.pair_s][q][i];
      double gi[2]; ugrad(g, D.face->grad[R.pair_s][q][i], gi);
      double u_s = 0.0, gus[2] = {0.0, 0.0};
      for (int j = 0; j < nk; ++j) { double gj[2]; ugrad(g, D.face->grad[R.pair_s][q][j], gj); u_s += xs[j] * D.face->val[R.pair_s][q][j]; gus[0] += xs[j] * gj[0]; gus[1] += xs[j] * gj[1]; }
      double xg[2] = {F.pa[0] + tq * F.d[0], F.pa[1] + tq * F.d[1]};
      if (nb < 0) {
        if (R.bc == 1) acc += eval_function(D.P.j, 2, xg, 0.0) * phi_i * factor;
        else if (R.bc == 0) {
          const double gval = eval_function(D.P.g, 2, xg, 0.0);
          acc += -(Ans[0] * gus[0] + Ans[1] * gus[1]) * factor * phi_i;
          acc += (u_s - gval) * factor * theta * (gi[0] * Ans[0] + gi[1] * Ans[1]);
          acc += R.penalty * (u_s - gval) * factor * phi_i;
        }
        continue;
      }
      double u_n = 0.0, gun[2] = {0.0, 0.0};
      for (int j = 0; j < nk; ++j) { double gj[2]; ugrad(gn, D.face->grad[R.pair_n][q][j], gj); u_n += xn[j] * D.face->val[R.pair_n][q][j]; gun[0] += xn[j] * gj[0]; gun[1] += xn[j] * gj[1]; }
      const double term2 = -(R.omega_s * (Ans[0] * gus[0] + Ans[1] * gus[1]) + R.omega_n * (Ann[0] * gun[0] + Ann[1] * gun[1])) * factor;
      const double term3 = (u_s - u_n) * factor;
      const double term4 = R.penalty * (u_s - u_n) * factor;
      acc += term2 * phi_i + term3 * theta * R.omega_s * (gi[0] * Ans[0] + gi[1] * Ans[1]) + term4 * phi_i;
    }
  }
  atomicAdd(&r[lm[i]], weight * acc);
}

// ---- StokesDarcyCouplingOperator (test/stokesdarcycouplingoperator.hh:104-245) -------------------------------------------------------------
struct USd { mdb_stokesdarcy_params P; const FeFace* vf; const FeFace* df; int vsize, dsize, qnl; int jac_mode; };   // qnl: local size of the Stokes subproblem (15)
struct USdFace { int pair_s, pair_n; UFaceGeo F; UTri gn; };

__device__ inline USdFace usd_face(const UM& m, int e, int f, int nb)
{
  USdFace S;
  const UTri gs = utri(m, e);
  S.F = uface(gs, f);
  S.gn = utri(m, nb);
  S.pair_s = 3 * S.F.la + S.F.lb;
  S.pair_n = uface_outside_pair(m, e, f, nb);
  return S;
}
// alpha_coupling, statement by statement; xs: the 2 * vsize velocity coefficients, xn: the head coefficients.  The pressure block
// receives nothing (:219-242).
__device__ void usd_alpha(const USd& C, const USdFace& S, const double* xs, const double* xn, double* rs, double* rn)
{
  const int vsize = C.vsize, dsize = C.dsize;
  const mdb_stokesdarcy_params& P = C.P;
  double tracePi = 0.0;
  for (int i = 0; i < 2; ++i) tracePi += P.kabs;
  tracePi *= P.viscosity / P.gravity / P.density;                                               // :172-175
  const double* n = S.F.n;
  for (int i = 0; i < 2 * vsize; ++i) rs[i] = 0.0;
  for (int i = 0; i < dsize; ++i) rn[i] = 0.0;
  for (int q = 0; q < C.df->nq; ++q) {
    const double tq = C.df->t[q];
    const double pos1 = S.F.pa[1] + tq * S.F.d[1];
    const double factor = C.df->w[q] * S.F.ie;
    const double* v = C.vf->val[S.pair_s][q];
    const double* psi = C.df->val[S.pair_n][q];
    double phi = 0.0, gradphi[2] = {0.0, 0.0};
    const int ngrad = P.quirk ? (vsize < dsize ? vsize : dsize) : dsize;                        // :194-195 loops i < vsize
    for (int i = 0; i < dsize; ++i) {
      double gp[2] = {0.0, 0.0};
      if (i < ngrad) ugrad(S.gn, C.df->grad[S.pair_n][q][i], gp);
      phi += xn[i] * psi[i];
      gradphi[0] += xn[i] * gp[0]; gradphi[1] += xn[i] * gp[1];
    }
    double u[2] = {0.0, 0.0};
    for (int d = 0; d < 2; ++d) for (int i = 0; i < vsize; ++i) u[d] += xs[vsize * d + i] * v[i];
    const double un = u[0] * n[0] + u[1] * n[1];
    for (int i = 0; i < dsize; ++i) rn[i] += -P.gamma * P.porosity * un * psi[i] * factor;      // :219
    double tang[2] = {P.kabs * gradphi[0] + 0.0 * gradphi[1], 0.0 * gradphi[0] + P.kabs * gradphi[1]};
    tang[0] /= P.porosity; tang[1] /= P.porosity;
    tang[0] += u[0]; tang[1] += u[1];
    const double tn = tang[0] * n[0] + tang[1] * n[1];
    tang[0] -= n[0] * tn; tang[1] -= n[1] * tn;                                                 // project into the tangential plane
    for (int d = 0; d < 2; ++d) {
      for (int i = 0; i < vsize; ++i) rs[vsize * d + i] += -P.density * P.gravity * (phi - pos1) * v[i] * n[d] * factor;                  // :235
      for (int i = 0; i < vsize; ++i) rs[vsize * d + i] += P.alpha * sqrt(2.0) / sqrt(tracePi) * tang[d] * v[i] * factor;                 // :240
    }
  }
}
__global__ void k_ufe_sd_residual(UM m, USd C, const int32_t* __restrict__ lmapL, const int32_t* __restrict__ lmapR, const int32_t* fcell,
                                  const int8_t* fface, int64_t nf, double weight, const double* __restrict__ x, double* r)
{
  const int64_t fi = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (fi >= nf) return;
  const int e = fcell[fi], f = fface[fi], nb = m.nbr[3 * e + f];
  const USdFace S = usd_face(m, e, f, nb);
  const int32_t* ls = lmapL + (int64_t)e * C.qnl; const int32_t* ln = lmapR + (int64_t)nb * C.dsize;
  double xs[12], xn[UFE_MAXN], rs[12], rn[UFE_MAXN];
  for (int i = 0; i < 2 * C.vsize; ++i) xs[i] = x[ls[i]];
  for (int i = 0; i < C.dsize; ++i) xn[i] = x[ln[i]];
  usd_alpha(C, S, xs, xn, rs, rn);
  for (int i = 0; i < 2 * C.vsize; ++i) atomicAdd(&r[ls[i]], weight * rs[i]);
  for (int i = 0; i < C.dsize; ++i) atomicAdd(&r[ln[i]], weight * rn[i]);
}
// NumericalJacobianCoupling (couplingutilities.hh:75-135): thread (face, j) evaluates the base line and the residual with coefficient j
// moved by delta = epsilon (1 + |u_j|), and adds column j.  Perturbing a pressure coefficient changes nothing (up = down bit for
// bit, entries exactly 0), so those columns are not evaluated.
__global__ void k_ufe_sd_jacobian(UM m, USd C, const int32_t* __restrict__ lmapL, const int32_t* __restrict__ lmapR, const int32_t* fcell,
                                  const int8_t* fface, int64_t nf, double weight, const double* __restrict__ x, const int64_t* __restrict__ rowptr,
                                  const int32_t* __restrict__ colidx, double* val)
{
  const int nv2 = 2 * C.vsize, ncol = nv2 + C.dsize;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= nf * ncol) return;
  const int64_t fi = idx / ncol; const int j = (int)(idx % ncol);
  const int e = fcell[fi], f = fface[fi], nb = m.nbr[3 * e + f];
  const USdFace S = usd_face(m, e, f, nb);
  const int32_t* ls = lmapL + (int64_t)e * C.qnl; const int32_t* ln = lmapR + (int64_t)nb * C.dsize;
  const bool analytic = C.jac_mode == MDB_JAC_ANALYTIC;
  double xs[12], xn[UFE_MAXN], down_s[12], down_n[UFE_MAXN], up_s[12], up_n[UFE_MAXN];
  for (int i = 0; i < nv2; ++i) xs[i] = analytic ? 0.0 : x[ls[i]];
  for (int i = 0; i < C.dsize; ++i) xn[i] = analytic ? 0.0 : x[ln[i]];
  usd_alpha(C, S, xs, xn, down_s, down_n);
  double* uj = j < nv2 ? &xs[j] : &xn[j - nv2];
  const double delta = analytic ? 1.0 : C.P.epsilon * (1.0 + fabs(*uj));
  *uj += delta;
  usd_alpha(C, S, xs, xn, up_s, up_n);
  const int32_t col = j < nv2 ? ls[j] : ln[j - nv2];
  for (int i = 0; i < ncol; ++i) {
    const int32_t row = i < nv2 ? ls[i] : ln[i - nv2];
    const double entry = i < nv2 ? (up_s[i] - down_s[i]) / delta : (up_n[i - nv2] - down_n[i - nv2]) / delta;
    const int64_t p0 = rowptr[row];
    const int pos = ufe_find(colidx + p0, (int)(rowptr[row + 1] - p0), col);
    if (pos >= 0) atomicAdd(&val[p0 + pos], weight * entry);
  }
}

__global__ void k_ufe_dirichlet_rows(const int32_t* conlist, int64_t ncon, const int64_t* rowptr, const int32_t* diag, double* val)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ncon) return;
  const int64_t row = conlist[i];
  const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
  for (int64_t p = p0; p < p1; ++p) val[p] = 0.0;
  if (diag[row] >= 0) val[p0 + diag[row]] = 1.0;
}

template <typename T> void ufree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

UTh th_view(Ctx* c, const SubProblemDesc& d)
{
  UTh T; T.P = d.th;
  T.v2 = vol_table(c, MDB_FE_P2, d.th.quad_order); T.p1 = vol_table(c, MDB_FE_P1, d.th.quad_order); T.v2f = face_table(c, MDB_FE_P2, d.th.quad_order);
  return T;
}
UDg dg_view(Ctx* c, const SubProblemDesc& d)
{
  const int fe = c->children[d.children[0]].fe_kind;
  UDg D; D.P = d.darcy; D.k = fe_order(fe); D.nk = fe_size(fe);
  const int order = d.darcy.intorderadd + 2 * D.k;
  D.vol = vol_table(c, fe, order); D.face = face_table(c, fe, order);
  return D;
}
USd sd_view(Ctx* c, const CouplingDesc& cp)
{
  const SubProblemDesc& L = c->sps[cp.local_sp]; const SubProblemDesc& R = c->sps[cp.remote_sp];
  const int fev = c->children[L.children[0]].fe_kind, fed = c->children[R.children[0]].fe_kind;
  USd C; C.P = cp.sd; C.vsize = fe_size(fev); C.dsize = fe_size(fed); C.qnl = L.nloc; C.jac_mode = cp.jac_mode;
  const int qorder = 2 * std::max(fe_order(fev), fe_order(fed));                                   // :159-160
  C.vf = face_table(c, fev, qorder); C.df = face_table(c, fed, qorder);
  return C;
}

} // anonymous namespace

// ---- host entry points ------------------------------------------------------------------------------------------------------------------------
void ufe_free(Ctx* c)
{
  if (c->um && c->um->ufe_tables) { delete (UfeTables*)c->um->ufe_tables; c->um->ufe_tables = nullptr; }
  for (SubProblemDesc& s : c->sps) ufree(s.d_lmap);
  for (Child& ch : c->children) { ufree(ch.d_edge2dof); ufree(ch.d_cell2dof); }
}

void ufe_build_dofmap(Ctx* c)
{
  const UM m = um_view(c);
  const int64_t ned = c->um->nedges;
  const int64_t nmax = std::max(std::max(m.nv, ned), m.ne);
  int32_t* d_vf = dalloc<int32_t>(m.nv); int32_t* d_ef = dalloc<int32_t>(ned); int32_t* d_cf = dalloc<int32_t>(m.ne);
  int32_t* d_scan = dalloc<int32_t>(nmax);
  int64_t offset = 0;
  for (Child& ch : c->children) {
    ch.um_nloc = fe_size(ch.fe_kind);
    ch.nlat = m.nv; ch.lat_n[0] = (int)m.nv; ch.lat_n[1] = ch.lat_n[2] = 1;
    ufree(ch.d_lat2dof); ufree(ch.d_dof2lat); ufree(ch.d_edge2dof); ufree(ch.d_cell2dof);
    MDB_CUDA(cudaMemsetAsync(d_vf, 0, m.nv * 4, c->stream)); MDB_CUDA(cudaMemsetAsync(d_ef, 0, ned * 4, c->stream));
    MDB_LAUNCH(c, k_ufe_mark, cdiv(m.ne, 256), 256, 0, m, ch.subdomain, d_vf, d_ef, d_cf);
    ch.d_lat2dof = dalloc<int32_t>(m.nv); ch.d_edge2dof = dalloc<int32_t>(ned); ch.d_cell2dof = dalloc<int32_t>(m.ne);
    const int64_t nvs = scan_i32(c, d_vf, d_scan, m.nv);
    MDB_LAUNCH(c, k_ufe_assign, cdiv(m.nv, 256), 256, 0, m.nv, d_vf, d_scan, ch.d_lat2dof);
    const int64_t nes = scan_i32(c, d_ef, d_scan, ned);
    MDB_LAUNCH(c, k_ufe_assign, cdiv(ned, 256), 256, 0, ned, d_ef, d_scan, ch.d_edge2dof);
    const int64_t ncs = scan_i32(c, d_cf, d_scan, m.ne);
    MDB_LAUNCH(c, k_ufe_assign, cdiv(m.ne, 256), 256, 0, m.ne, d_cf, d_scan, ch.d_cell2dof);
    ch.nvs = nvs;
    ch.size = ch.fe_kind == MDB_FE_P1 ? nvs : ch.fe_kind == MDB_FE_P2 ? nvs + nes : ncs * ch.um_nloc;
    ch.offset = offset; offset += ch.size;
  }
  MDB_REQUIRE(offset < (int64_t)2000000000, MDB_EINVAL, "unstructured mesh: more than 2e9 DOFs per context");
  c->ndof = offset;
  ufree(c->d_constrained);
  c->d_constrained = dalloc<uint8_t>(c->ndof);
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  // the local function space of every subproblem, once (SubProblemLocalFunctionSpace, a5)
  for (SubProblemDesc& s : c->sps) {
    const USp S = sp_view(c, s);
    MDB_REQUIRE(S.nloc <= UFE_MAXLOC, MDB_EINVAL, "generic elements: a subproblem's local space has at most 16 functions");
    s.nloc = S.nloc;
    ufree(s.d_lmap);
    s.d_lmap = dalloc<int32_t>(m.ne * (int64_t)S.nloc);
    MDB_LAUNCH(c, k_ufe_lmap, cdiv(m.ne, 256), 256, 0, m, S, s.d_lmap);
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_vf); cudaFree(d_ef); cudaFree(d_cf); cudaFree(d_scan);
}

// parity export: the MultiDomainLocalFunctionSpace of every cell, children concatenated (multidomainlocalfunctionspace.hh:348-358)
void ufe_export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs)
{
  const UMesh* u = c->um;
  std::vector<int32_t> conn(3 * u->ne), ced(3 * u->ne); std::vector<uint8_t> sd(u->ne);
  MDB_CUDA(cudaMemcpy(conn.data(), u->d_conn, 3 * u->ne * 4, cudaMemcpyDeviceToHost));
  MDB_CUDA(cudaMemcpy(ced.data(), u->d_cell_edge, 3 * u->ne * 4, cudaMemcpyDeviceToHost));
  MDB_CUDA(cudaMemcpy(sd.data(), c->d_cell_sd, u->ne, cudaMemcpyDeviceToHost));
  const size_t nch = c->children.size();
  std::vector<std::vector<int32_t>> v2d(nch), e2d(nch), c2d(nch);
  for (size_t k = 0; k < nch; ++k) {
    v2d[k].resize(u->nv); e2d[k].resize(u->nedges); c2d[k].resize(u->ne);
    MDB_CUDA(cudaMemcpy(v2d[k].data(), c->children[k].d_lat2dof, u->nv * 4, cudaMemcpyDeviceToHost));
    MDB_CUDA(cudaMemcpy(e2d[k].data(), c->children[k].d_edge2dof, u->nedges * 4, cudaMemcpyDeviceToHost));
    MDB_CUDA(cudaMemcpy(c2d[k].data(), c->children[k].d_cell2dof, u->ne * 4, cudaMemcpyDeviceToHost));
  }
  int64_t p = 0;
  for (int64_t e = 0; e < u->ne; ++e) {
    cell_ptr[e] = p;
    for (size_t k = 0; k < nch; ++k) {
      const Child& ch = c->children[k];
      if (!(ch.subdomain < 0 || ((sd[e] >> ch.subdomain) & 1))) continue;
      if (cell_dofs) {
        const int32_t* v = &conn[3 * e]; const int32_t* g = &ced[3 * e];
        if (ch.fe_kind == MDB_FE_P1) for (int i = 0; i < 3; ++i) cell_dofs[p + i] = ch.offset + v2d[k][v[i]];
        else if (ch.fe_kind == MDB_FE_P2) {
          const int64_t vv[3] = {v2d[k][v[0]], v2d[k][v[1]], v2d[k][v[2]]};
          const int64_t ee[3] = {ch.nvs + e2d[k][g[0]], ch.nvs + e2d[k][g[1]], ch.nvs + e2d[k][g[2]]};
          const int64_t loc[6] = {vv[0], ee[0], vv[1], ee[1], ee[2], vv[2]};
          for (int i = 0; i < 6; ++i) cell_dofs[p + i] = ch.offset + loc[i];
        } else for (int i = 0; i < ch.um_nloc; ++i) cell_dofs[p + i] = ch.offset + (int64_t)c2d[k][e] * ch.um_nloc + i;
      }
      p += ch.um_nloc;
    }
  }
  cell_ptr[u->ne] = p;
}

void ufe_assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n)
{
  MDB_REQUIRE(n <= 8, MDB_EINVAL, "unstructured mesh: at most 8 subproblems per call");
  const UM m = um_view(c);
  UCon B; B.n = n;
  for (int i = 0; i < n; ++i) {
    const SubProblemDesc& d = c->sps[sps[i]];
    B.lmap[i] = d.d_lmap; B.nloc[i] = d.nloc; B.bc[i] = bcs[i];
    // [UPSTREAM] StokesVelocityDirichletConstraints x StokesPressureDirichletConstraints: the velocity children only
    B.nkids[i] = d.op_id == MDB_OP_TAYLORHOOD_STOKES ? 2 : (int)d.children.size();
    for (int k = 0; k < B.nkids[i]; ++k) B.kid_fe[i][k] = c->children[d.children[k]].fe_kind;
  }
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  MDB_LAUNCH(c, k_ufe_constraints, cdiv(3 * m.ne, 256), 256, 0, m, B, c->d_constrained);
}

void ufe_build_facelists(Ctx* c, GridOp& go)
{
  const UM m = um_view(c);
  const int64_t n = 3 * m.ne;
  int32_t* d_flag = dalloc<int32_t>(n); int32_t* d_scan = dalloc<int32_t>(n);
  go.faces.clear();
  for (int q : go.cps) {
    const CouplingDesc& cp = c->cps[q];
    const SubProblemDesc& L = c->sps[cp.local_sp]; const SubProblemDesc& R = c->sps[cp.remote_sp];
    MDB_LAUNCH(c, k_ufe_face_flags, cdiv(n, 256), 256, 0, m, L.d_lmap, L.nloc, R.d_lmap, R.nloc, d_flag);
    FaceList fl; fl.n = scan_i32(c, d_flag, d_scan, n);
    fl.d_cell = dalloc<int32_t>(fl.n); fl.d_face = dalloc<int8_t>(fl.n);
    MDB_LAUNCH(c, k_ufe_face_compact, cdiv(n, 256), 256, 0, n, d_flag, d_scan, fl.d_cell, fl.d_face);
    go.faces.push_back(fl);
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_flag); cudaFree(d_scan);
}

void ufe_build_pattern(Ctx* c, Matrix& M)
{
  const UM m = um_view(c);
  std::vector<int> sps; std::vector<std::pair<int, const FaceList*>> cps;
  for (int g : M.gos) {
    for (int s : c->gos[g].sps) if (std::find(sps.begin(), sps.end(), s) == sps.end()) sps.push_back(s);
    for (size_t i = 0; i < c->gos[g].cps.size(); ++i) {
      const int q = c->gos[g].cps[i];
      bool seen = false; for (auto& p : cps) seen = seen || p.first == q;
      if (!seen) cps.push_back({q, &c->gos[g].faces[i]});
    }
  }
  // segment sizes
  int64_t nkeys = 0;
  std::vector<int64_t> sp_base, cp_base;
  for (int s : sps) { const SubProblemDesc& d = c->sps[s]; sp_base.push_back(nkeys); nkeys += m.ne * (int64_t)d.nloc * d.nloc * (d.op_id == MDB_OP_DARCY_DG ? 4 : 1); }
  for (auto& p : cps) { const CouplingDesc& cp = c->cps[p.first]; cp_base.push_back(nkeys); nkeys += p.second->n * 2 * (int64_t)c->sps[cp.local_sp].nloc * c->sps[cp.remote_sp].nloc; }
  MDB_REQUIRE(nkeys < ((int64_t)1 << 31), MDB_EINVAL, "generic elements: more than 2^31 pattern keys (mesh too large for one pass)");
  const int64_t n = c->ndof;
  M.d_rowptr = dalloc<int64_t>(n + 1);
  MDB_CUDA(cudaMemsetAsync(M.d_rowptr, 0, (n + 1) * sizeof(int64_t), c->stream));
  int64_t nnz = 0;
  uint64_t* d_keys = dalloc<uint64_t>(nkeys); uint64_t* d_sorted = dalloc<uint64_t>(nkeys); uint64_t* d_uniq = nullptr;
  if (nkeys > 0) {
    for (size_t i = 0; i < sps.size(); ++i) {
      const SubProblemDesc& d = c->sps[sps[i]];
      MDB_LAUNCH(c, k_ufe_keys_volume, cdiv(m.ne * (int64_t)d.nloc, 128), 128, 0, m, d.d_lmap, d.nloc, d.op_id == MDB_OP_DARCY_DG ? 1 : 0, d_keys + sp_base[i]);
    }
    for (size_t i = 0; i < cps.size(); ++i) {
      const CouplingDesc& cp = c->cps[cps[i].first]; const FaceList& fl = *cps[i].second;
      const SubProblemDesc& L = c->sps[cp.local_sp]; const SubProblemDesc& R = c->sps[cp.remote_sp];
      if (fl.n > 0) MDB_LAUNCH(c, k_ufe_keys_coupling, cdiv(fl.n * (L.nloc + R.nloc), 128), 128, 0, m, L.d_lmap, L.nloc, R.d_lmap, R.nloc, fl.d_cell, fl.d_face, fl.n, d_keys + cp_base[i]);
    }
    // segmented sort + unique of (row, column): rows ascending, columns ascending inside a row = the BCRS order
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, d_keys, d_sorted, (int)nkeys, 0, 64, c->stream);
    void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
    MDB_CUDA(cub::DeviceRadixSort::SortKeys(d_t, tb, d_keys, d_sorted, (int)nkeys, 0, 64, c->stream)); c->launches++;
    cudaFree(d_t);
    int* d_num = dalloc<int>(1);
    tb = 0;
    cub::DeviceSelect::Unique(nullptr, tb, d_sorted, d_keys, d_num, (int)nkeys, c->stream);
    MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
    MDB_CUDA(cub::DeviceSelect::Unique(d_t, tb, d_sorted, d_keys, d_num, (int)nkeys, c->stream)); c->launches++;
    int num = 0; uint64_t last = 0;
    MDB_CUDA(cudaMemcpyAsync(&num, d_num, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t); cudaFree(d_num);
    if (num > 0) MDB_CUDA(cudaMemcpy(&last, d_keys + num - 1, 8, cudaMemcpyDeviceToHost));
    nnz = (num > 0 && last == UFE_NOKEY) ? num - 1 : num;
    d_uniq = d_keys;
  }
  M.nnz = nnz;
  M.d_colidx = dalloc<int32_t>(nnz); M.d_val = dalloc<double>(nnz); M.d_diag = dalloc<int32_t>(n);
  MDB_CUDA(cudaMemsetAsync(M.d_val, 0, nnz * sizeof(double), c->stream));
  if (nnz > 0) MDB_LAUNCH(c, k_ufe_csr, cdiv(nnz, 256), 256, 0, d_uniq, nnz, n, M.d_rowptr, M.d_colidx);
  int32_t* d_len = dalloc<int32_t>(n);
  MDB_LAUNCH(c, k_ufe_diag, cdiv(n, 256), 256, 0, n, M.d_rowptr, M.d_colidx, M.d_diag, d_len);
  int32_t maxrow = 0;
  {
    size_t tb2 = 0; int32_t* d_max = dalloc<int32_t>(1);
    cub::DeviceReduce::Max(nullptr, tb2, d_len, d_max, (int)n, c->stream);
    void* d_t2 = nullptr; MDB_CUDA(cudaMalloc(&d_t2, tb2 ? tb2 : 1));
    MDB_CUDA(cub::DeviceReduce::Max(d_t2, tb2, d_len, d_max, (int)n, c->stream)); c->launches++;
    MDB_CUDA(cudaMemcpyAsync(&maxrow, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t2); cudaFree(d_max);
  }
  M.max_row = maxrow;
  for (size_t k = 0; k < c->children.size(); ++k) { M.child_begin[k] = c->children[k].offset; M.child_end[k] = c->children[k].offset + c->children[k].size; }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_keys); cudaFree(d_sorted); cudaFree(d_len);
}

void ufe_residual(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r)
{
  const UM m = um_view(c);
  c->phase_begin("residual_volume");
  for (int s : go.sps) {
    const SubProblemDesc& d = c->sps[s];
    if (d.op_id == MDB_OP_TAYLORHOOD_STOKES) MDB_LAUNCH(c, k_ufe_th_residual, cdiv(m.ne * 15, 128), 128, 0, m, th_view(c, d), d.d_lmap, weight, d_x, d_r);
    else { const UDg D = dg_view(c, d); MDB_LAUNCH(c, k_ufe_dg_residual, cdiv(m.ne * (int64_t)D.nk, 128), 128, 0, m, D, d.d_lmap, weight, d_x, d_r); }
  }
  c->phase_end();
  c->phase_begin("residual_coupling");
  for (size_t i = 0; i < go.cps.size(); ++i) {
    const CouplingDesc& cp = c->cps[go.cps[i]]; const FaceList& fl = go.faces[i];
    if (fl.n > 0) MDB_LAUNCH(c, k_ufe_sd_residual, cdiv(fl.n, 64), 64, 0, m, sd_view(c, cp), c->sps[cp.local_sp].d_lmap, c->sps[cp.remote_sp].d_lmap, fl.d_cell, fl.d_face, fl.n, weight, d_x, d_r);
  }
  constrain_residual(c, d_r);
  c->phase_end();
}

void ufe_jacobian(Ctx* c, const GridOp& go, Matrix& M, double weight, const double* d_x, bool overwrite)
{
  const UM m = um_view(c);
  if (overwrite) MDB_CUDA(cudaMemsetAsync(M.d_val, 0, M.nnz * sizeof(double), c->stream));
  c->phase_begin("jacobian_rows");
  for (int s : go.sps) {
    const SubProblemDesc& d = c->sps[s];
    if (d.op_id == MDB_OP_TAYLORHOOD_STOKES) MDB_LAUNCH(c, k_ufe_th_jacobian, cdiv(m.ne * 15, 128), 128, 0, m, th_view(c, d), d.d_lmap, weight, M.d_rowptr, M.d_colidx, M.d_val);
    else { const UDg D = dg_view(c, d); MDB_LAUNCH(c, k_ufe_dg_jacobian, cdiv(m.ne * (int64_t)D.nk, 128), 128, 0, m, D, d.d_lmap, weight, M.d_rowptr, M.d_colidx, M.d_val); }
  }
  c->phase_end();
  c->phase_begin("jacobian_coupling");
  for (size_t i = 0; i < go.cps.size(); ++i) {
    const CouplingDesc& cp = c->cps[go.cps[i]]; const FaceList& fl = go.faces[i];
    if (fl.n == 0) continue;
    const USd C = sd_view(c, cp);
    MDB_LAUNCH(c, k_ufe_sd_jacobian, cdiv(fl.n * (2 * C.vsize + C.dsize), 64), 64, 0, m, C, c->sps[cp.local_sp].d_lmap, c->sps[cp.remote_sp].d_lmap, fl.d_cell, fl.d_face, fl.n,
               weight, d_x, M.d_rowptr, M.d_colidx, M.d_val);
  }
  c->phase_end();
  c->phase_begin("jacobian_dirichlet");
  if (c->ncon > 0) MDB_LAUNCH(c, k_ufe_dirichlet_rows, cdiv(c->ncon, 128), 128, 0, c->d_conlist, c->ncon, M.d_rowptr, M.d_diag, M.d_val);
  c->phase_end();
}

} // namespace mdb
