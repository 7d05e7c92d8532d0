// unstructured.cu — the assembly path on an unstructured conforming triangle mesh with P1 spaces: the mesh class the reference's
// Stokes-Darcy drivers read with GmshReader<UGGrid<2>> (test/stokesdarcy2.cc:575-604) and the north star's "refined gmsh meshes".
//
// Same design as the structured path, without the lattice:
//   DOF map      vertex flags -> scan -> assign per child (ascending vertex index of the child's subdomain,
//                multidomaingridfunctionspace.hh:308-369 with [UPSTREAM] (ii),(iii))
//   pattern      one thread per row: union of the links of the incident cells (FullVolumePattern) and of the coupling faces
//                of those cells (couplingutilities.hh:35-47), sorted-unique insertion, count + fill
//   Jacobian     owner-computes rows: the thread of a row sums its incident cells in ascending cell order (the reference's
//                accumulation order) and writes every stored entry once; coupling faces add by atomics afterwards
//   residual     owner-computes rows for the volume terms; boundary and coupling faces add by atomics
// Geometry is per element (affine): gradients of the P1 shape functions and the integration element are recomputed from the
// vertex coordinates where they are used (3 coordinates pairs per cell = 48 B against 72 B of matrix row written).
// Compiled with --fmad=false: the finite-difference coupling Jacobian must see the reference's operation order.
#include "common.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace mdb {

namespace {

constexpr int UM_MAXROW = 128;     // column capacity of the per-thread sorted set while a row's pattern is enumerated

struct UMDev {
  int64_t nv, ne;
  const double* xy; const int32_t* conn; const int32_t* nbr; const int32_t* v2c_ptr; const int32_t* v2c; const uint8_t* sd;
};
struct UMSpace { int nchild; int64_t offset[MDB_MAX_CHILDREN]; const int32_t* v2d[MDB_MAX_CHILDREN]; const int32_t* d2v[MDB_MAX_CHILDREN]; int64_t size[MDB_MAX_CHILDREN]; };
struct UMSp { int op_id, child, cond_kind; uint32_t mask; int quad_order; double wsum; mdb_poisson_params P; };   // wsum: sum of the triangle rule's weights, in rule order
struct UMCp { int op_id, childL, childR, kindL, kindR; uint32_t maskL, maskR; int jac_mode, quad_order; double intensity; };
struct UMOps { int nsp, ncp; UMSp sp[8]; UMCp cp[4]; };

double tri_weight_sum(int order);

UMDev um_dev(const Ctx* c)
{
  const UMesh* u = c->um;
  UMDev m; m.nv = u->nv; m.ne = u->ne; m.xy = u->d_xy; m.conn = u->d_conn; m.nbr = u->d_nbr; m.v2c_ptr = u->d_v2c_ptr; m.v2c = u->d_v2c; m.sd = c->d_cell_sd;
  return m;
}
UMSpace um_space(const Ctx* c)
{
  UMSpace S; S.nchild = (int)c->children.size();
  for (int i = 0; i < S.nchild; ++i) { S.offset[i] = c->children[i].offset; S.v2d[i] = c->children[i].d_lat2dof; S.d2v[i] = c->children[i].d_dof2lat; S.size[i] = c->children[i].size; }
  return S;
}
UMOps um_ops(const Ctx* c, const std::vector<int>& sps, const std::vector<int>& cps)
{
  UMOps O; O.nsp = 0; O.ncp = 0;
  MDB_REQUIRE(sps.size() <= 8 && cps.size() <= 4, MDB_EINVAL, "unstructured mesh: at most 8 subproblems and 4 couplings per grid operator");
  for (int s : sps) {
    const SubProblemDesc& d = c->sps[s];
    UMSp& o = O.sp[O.nsp++];
    o.op_id = d.op_id; o.child = d.children[0]; o.cond_kind = d.cond_kind; o.mask = d.mask; o.quad_order = d.quad_order(); o.P = d.poisson;
    o.wsum = tri_weight_sum(o.quad_order);
  }
  for (int q : cps) {
    const CouplingDesc& d = c->cps[q];
    const SubProblemDesc& L = c->sps[d.local_sp]; const SubProblemDesc& R = c->sps[d.remote_sp];
    UMCp& o = O.cp[O.ncp++];
    o.op_id = d.op_id; o.childL = L.children[0]; o.childR = R.children[0]; o.kindL = L.cond_kind; o.kindR = R.cond_kind; o.maskL = L.mask; o.maskR = R.mask;
    o.jac_mode = d.jac_mode;
    o.intensity = d.op_id == MDB_OP_CVCF_COUPLING ? d.cvcf.intensity : d.propflow.intensity;
    o.quad_order = d.op_id == MDB_OP_CVCF_COUPLING ? d.cvcf.quad_order : 4;     // ProportionalFlowCoupling: intorder fixed to 4 (test/proportionalflowcoupling.hh:43)
  }
  return O;
}

// ---- triangle geometry and rules ---------------------------------------------------------------------------------------
// Reference triangle (0,0),(1,0),(0,1) [UPSTREAM dune-geometry]; x = p0 + J xi, J = [p1 - p0 | p2 - p0].
struct Tri {
  double p[3][2];
  double t[2][2];      // J^{-T}
  double ie;           // integration element |det J|
};
__device__ inline Tri tri_geometry(const UMDev& m, int cell)
{
  Tri g;
  for (int i = 0; i < 3; ++i) { const int v = m.conn[3 * cell + i]; g.p[i][0] = m.xy[2 * v]; g.p[i][1] = m.xy[2 * v + 1]; }
  const double ax = g.p[1][0] - g.p[0][0], ay = g.p[1][1] - g.p[0][1];
  const double bx = g.p[2][0] - g.p[0][0], by = g.p[2][1] - g.p[0][1];
  const double det = ax * by - ay * bx;
  const double inv = 1.0 / det;                  // one division per cell (the reference divides four times: 1 ulp apart)
  g.t[0][0] = by * inv; g.t[0][1] = -(ay * inv);
  g.t[1][0] = -(bx * inv); g.t[1][1] = ax * inv;
  g.ie = fabs(det);
  return g;
}
// gradients of the P1 shape functions: J^{-T} applied to (-1,-1), (1,0), (0,1)
__device__ inline void tri_gradients(const Tri& g, double gp[3][2])
{
  for (int d = 0; d < 2; ++d) { gp[1][d] = g.t[d][0]; gp[2][d] = g.t[d][1]; gp[0][d] = -g.t[d][0] - g.t[d][1]; }
}
__device__ inline void p1_basis(double x, double y, double phi[3]) { phi[0] = 1.0 - x - y; phi[1] = x; phi[2] = y; }
// v[i] for a run-time i without putting v into local memory
__device__ inline double pick3(const double v0, const double v1, const double v2, int i) { return i == 0 ? v0 : (i == 1 ? v1 : v2); }

// [UPSTREAM] Dune::QuadratureRules<DF,2>::rule(simplex, order): published symmetric rules on the reference triangle (weights sum
// to 1/2): order <= 1 centroid; order 2 the 3-point interior rule; orders 3, 4 the 6-point degree-4 rule (Dunavant).
#define UM_TA 0.445948490915965
#define UM_TB 0.091576213509771
#define UM_WA (0.5 * 0.223381589678011)
#define UM_WB (0.5 * 0.109951743655322)
// rows 0 | 1-3 | 4-9: the 1-, 3- and 6-point rules (x, y, weight)
#define UM_TRI_TABLE {{1.0 / 3.0, 1.0 / 3.0, 0.5},                                                                                  \
                      {1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0}, {2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0},         \
                      {UM_TA, UM_TA, UM_WA}, {1.0 - 2.0 * UM_TA, UM_TA, UM_WA}, {UM_TA, 1.0 - 2.0 * UM_TA, UM_WA},                    \
                      {UM_TB, UM_TB, UM_WB}, {1.0 - 2.0 * UM_TB, UM_TB, UM_WB}, {UM_TB, 1.0 - 2.0 * UM_TB, UM_WB}}
__device__ __constant__ double c_tri[10][3] = UM_TRI_TABLE;
static const double h_tri[10][3] = UM_TRI_TABLE;
// number of points; *first = first row of the rule in the table
__host__ __device__ inline int tri_rule(int order, int* first)
{
  if (order <= 1) { *first = 0; return 1; }
  if (order == 2) { *first = 1; return 3; }
  *first = 4; return 6;
}
double tri_weight_sum(int order)
{
  int q0; const int nq = tri_rule(order, &q0);
  double s = 0.0;
  for (int q = 0; q < nq; ++q) s += h_tri[q0 + q][2];
  return s;
}
// Gauss-Legendre on [0,1], order p -> ceil((p+1)/2) points (<= 3 here)
__device__ inline int line_rule(int order, double x[3], double w[3])
{
  const int m = (order + 2) / 2;
  if (m <= 1) { x[0] = 0.5; w[0] = 1.0; return 1; }
  if (m == 2) { const double a = 0.5 / sqrt(3.0); x[0] = 0.5 - a; x[1] = 0.5 + a; w[0] = 0.5; w[1] = 0.5; return 2; }
  const double a = 0.5 * sqrt(3.0 / 5.0);
  x[0] = 0.5 - a; x[1] = 0.5; x[2] = 0.5 + a; w[0] = 5.0 / 18.0; w[1] = 8.0 / 18.0; w[2] = 5.0 / 18.0; return 3;
}
// local vertices of face f and the reference corners
__device__ __constant__ int c_face_v[3][2] = {{0, 1}, {0, 2}, {1, 2}};
__device__ __constant__ double c_ref[3][2] = {{0.0, 0.0}, {1.0, 0.0}, {0.0, 1.0}};

__device__ inline bool um_applies(const UMSp& s, uint32_t sd) { return cond_applies(s.cond_kind, s.mask, sd); }

__device__ inline int um_find(const int32_t* cols, int len, int32_t col)
{
  int lo = 0, hi = len - 1;
  while (lo <= hi) { const int mid = (lo + hi) >> 1; const int32_t v = cols[mid]; if (v == col) return mid; if (v < col) lo = mid + 1; else hi = mid - 1; }
  return -1;
}

// ---- DOF map --------------------------------------------------------------------------------------------------------------
__global__ void k_um_mark(UMDev m, int subdomain, int32_t* flag)
{
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= m.nv) return;
  int f = 0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const int cell = m.v2c[e] >> 2;
    if (subdomain < 0 || ((m.sd[cell] >> subdomain) & 1)) { f = 1; break; }
  }
  flag[v] = f;
}
__global__ void k_um_assign(int64_t nv, const int32_t* flag, const int32_t* scan, int32_t* v2d, int32_t* d2v)
{
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (flag[v]) { v2d[v] = scan[v]; d2v[scan[v]] = (int32_t)v; } else v2d[v] = -1;
}

// ---- constraints + interpolation ------------------------------------------------------------------------------------------
struct UMBcs { int n; UMSp sp[8]; mdb_bctype bc[8]; mdb_function fn[8]; };

// constraintsfunctors.hh:75-103: on a face whose outside cell is not in the subproblem (physical boundary or subdomain interface)
// the boundary constraints run; [UPSTREAM] ConformingDirichletConstraints asks the predicate at the face centre and constrains the
// coefficients attached to the face's sub-entities (P1: its two vertices).
__global__ void k_um_constraints(UMDev m, UMSpace S, UMBcs B, uint8_t* constrained)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m.ne) return;
  const int cell = (int)(i / 3), f = (int)(i % 3);
  const uint32_t sd = m.sd[cell];
  const int nb = m.nbr[3 * cell + f];
  const bool interior = nb >= 0;
  const int va = m.conn[3 * cell + c_face_v[f][0]], vb = m.conn[3 * cell + c_face_v[f][1]];
  double xg[2];
  for (int d = 0; d < 2; ++d) { const double pa = m.xy[2 * va + d], pb = m.xy[2 * vb + d]; xg[d] = pa + 0.5 * (pb - pa); }
  for (int q = 0; q < B.n; ++q) {
    const UMSp& s = B.sp[q];
    if (!um_applies(s, sd)) continue;
    if (interior && um_applies(s, m.sd[nb])) continue;
    if (eval_bctype(B.bc[q], 2, xg, !interior) != 0) continue;
    const int32_t da = S.v2d[s.child][va], db = S.v2d[s.child][vb];
    if (da >= 0) constrained[S.offset[s.child] + da] = 1;
    if (db >= 0) constrained[S.offset[s.child] + db] = 1;
  }
}

// interpolate.hh:35-73: cells in order, every subproblem that applies writes f at the cell's vertices; per DOF the last writer wins
__global__ void k_um_interpolate(UMDev m, UMSpace S, UMBcs B, int child, double t, double* x)
{
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= S.size[child]) return;
  const int v = S.d2v[child][d];
  const double xv[2] = {m.xy[2 * v], m.xy[2 * v + 1]};
  bool have = false; double val = 0.0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const uint32_t sd = m.sd[m.v2c[e] >> 2];
    for (int q = 0; q < B.n; ++q)
      if (B.sp[q].child == child && um_applies(B.sp[q], sd)) { val = eval_function(B.fn[q], 2, xv, t); have = true; }
  }
  if (have) x[S.offset[child] + d] = val;
}

// ---- coupling faces ---------------------------------------------------------------------------------------------------------
__global__ void k_um_face_flags(UMDev m, UMCp cp, int32_t* flag)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 3 * m.ne) return;
  const int nb = m.nbr[i];
  flag[i] = (nb >= 0 && cond_applies(cp.kindL, cp.maskL, m.sd[i / 3]) && cond_applies(cp.kindR, cp.maskR, m.sd[nb])) ? 1 : 0;
}
__global__ void k_um_face_compact(int64_t n, const int32_t* flag, const int32_t* scan, int32_t* cell, int8_t* face)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  cell[scan[i]] = (int32_t)(i / 3); face[scan[i]] = (int8_t)(i % 3);
}

// ---- pattern ----------------------------------------------------------------------------------------------------------------
__device__ inline void um_insert(int32_t* cols, int& n, int32_t col, int* overflow)
{
  int pos = n;
  while (pos > 0 && cols[pos - 1] > col) --pos;
  if (pos > 0 && cols[pos - 1] == col) return;
  if (n >= UM_MAXROW) { *overflow = 1; return; }
  for (int k = n; k > pos; --k) cols[k] = cols[k - 1];
  cols[pos] = col; ++n;
}
__device__ inline void um_insert_cell(const UMDev& m, const UMSpace& S, int child, int cell, int32_t* cols, int& n, int* overflow)
{
  for (int j = 0; j < 3; ++j) {
    const int32_t d = S.v2d[child][m.conn[3 * cell + j]];
    if (d >= 0) um_insert(cols, n, (int32_t)(S.offset[child] + d), overflow);
  }
}
// columns of the row of vertex v in child block `child`
__device__ int um_row_columns(const UMDev& m, const UMSpace& S, const UMOps& O, int child, int v, int32_t* cols, int* overflow)
{
  int n = 0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const int cell = m.v2c[e] >> 2;
    const uint32_t sd = m.sd[cell];
    for (int s = 0; s < O.nsp; ++s)
      if (O.sp[s].child == child && um_applies(O.sp[s], sd)) um_insert_cell(m, S, child, cell, cols, n, overflow);
    for (int q = 0; q < O.ncp; ++q) {
      const UMCp& cp = O.cp[q];
      // pattern_coupling links lfsv_s x lfsu_n and lfsv_n x lfsu_s of every face the coupling fires on (couplingutilities.hh:35-47)
      if (cp.childL == child && cond_applies(cp.kindL, cp.maskL, sd))
        for (int f = 0; f < 3; ++f) { const int nb = m.nbr[3 * cell + f]; if (nb >= 0 && cond_applies(cp.kindR, cp.maskR, m.sd[nb])) um_insert_cell(m, S, cp.childR, nb, cols, n, overflow); }
      if (cp.childR == child && cond_applies(cp.kindR, cp.maskR, sd))
        for (int f = 0; f < 3; ++f) { const int nb = m.nbr[3 * cell + f]; if (nb >= 0 && cond_applies(cp.kindL, cp.maskL, m.sd[nb])) um_insert_cell(m, S, cp.childL, nb, cols, n, overflow); }
    }
  }
  return n;
}
__global__ void k_um_pattern_count(UMDev m, UMSpace S, UMOps O, int child, int32_t* count, int* overflow)
{
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= S.size[child]) return;
  int32_t cols[UM_MAXROW];
  count[S.offset[child] + d] = um_row_columns(m, S, O, child, S.d2v[child][d], cols, overflow);
}
// Also fills the slot table of the child block: for every (vertex, incident cell) entry e of v2c and every local vertex j of that
// cell, slot[3 e + j] = position inside the row of (child, vertex) of the column of (child, vertex j of the cell); 255 = not stored.
// The Jacobian kernel places the element-matrix rows with it: no column search and no DOF-map gather per assembly.
__global__ void k_um_pattern_fill(UMDev m, UMSpace S, UMOps O, int child, const int64_t* rowptr, int32_t* colidx, int32_t* diag, uint8_t* slot)
{
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= S.size[child]) return;
  int32_t cols[UM_MAXROW]; int ov = 0;
  const int64_t row = S.offset[child] + d;
  const int n = um_row_columns(m, S, O, child, S.d2v[child][d], cols, &ov);
  const int64_t p0 = rowptr[row];
  int dg = -1;
  for (int k = 0; k < n; ++k) { colidx[p0 + k] = cols[k]; if (cols[k] == (int32_t)row) dg = k; }
  diag[row] = dg;
  const int v = S.d2v[child][d];
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const int cell = m.v2c[e] >> 2;
    for (int j = 0; j < 3; ++j) {
      const int32_t dj = S.v2d[child][m.conn[3 * cell + j]];
      const int pos = dj >= 0 ? um_find(cols, n, (int32_t)(S.offset[child] + dj)) : -1;
      slot[3 * (int64_t)e + j] = (pos >= 0 && pos < 255) ? (uint8_t)pos : (uint8_t)255;
    }
  }
}
__global__ void k_um_widen(int64_t n, const int32_t* count, const int64_t* scan, int64_t* rowptr)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) rowptr[i] = scan[i];
  if (i == n - 1) rowptr[n] = scan[i] + count[i];
}
__global__ void k_um_to64(int64_t n, const int32_t* in, int64_t* out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// ---- volume terms ---------------------------------------------------------------------------------------------------------------
// Row i of the element matrix of subproblem s on a cell: Poisson ([UPSTREAM] PDELab::Poisson / test/instationarypoissonoperator.hh:52-77)
// or L2 ([UPSTREAM] PDELab::L2 / test/simpletimeoperator.hh:29-101), accumulated point by point with the weight as the reference's
// weighted local views do (localassembler.hh:214-255).
__device__ inline void um_matrix_row(const UMSp& s, const Tri& g, const double gp[3][2], int i, double weight, double a[3])
{
  if (s.op_id == MDB_OP_POISSON) {
    // the integrand is constant on the cell: sum_q weight (dot w_q |det J|) = weight (dot (sum_q w_q) |det J|) up to rounding
    const double vol = s.wsum * g.ie;
    const double gix = pick3(gp[0][0], gp[1][0], gp[2][0], i), giy = pick3(gp[0][1], gp[1][1], gp[2][1], i);
#pragma unroll
    for (int j = 0; j < 3; ++j) { double dot = 0.0; dot += gp[j][0] * gix; dot += gp[j][1] * giy; a[j] = weight * (dot * vol); }
    return;
  }
  int q0; const int nq = tri_rule(s.quad_order, &q0);
  a[0] = a[1] = a[2] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double factor = c_tri[q0 + q][2] * g.ie;
    double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
    const double phii = pick3(phi[0], phi[1], phi[2], i);
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] += weight * (phi[j] * phii * factor);
  }
}

__global__ void k_um_jacobian_rows(UMDev m, UMSpace S, UMOps O, int child, double weight, int overwrite, const int64_t* rowptr, const int32_t* colidx, double* val)
{
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= S.size[child]) return;
  const int v = S.d2v[child][d];
  const int64_t row = S.offset[child] + d;
  const int64_t p0 = rowptr[row]; const int len = (int)(rowptr[row + 1] - p0);
  if (overwrite) for (int k = 0; k < len; ++k) val[p0 + k] = 0.0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const int cell = m.v2c[e] >> 2, li = m.v2c[e] & 3;
    const uint32_t sd = m.sd[cell];
    bool any = false;
    for (int s = 0; s < O.nsp; ++s) any = any || (O.sp[s].child == child && um_applies(O.sp[s], sd));
    if (!any) continue;
    const Tri g = tri_geometry(m, cell);
    double gp[3][2]; tri_gradients(g, gp);
    int pos[3];
    for (int j = 0; j < 3; ++j) pos[j] = um_find(colidx + p0, len, (int32_t)(S.offset[child] + S.v2d[child][m.conn[3 * cell + j]]));
    for (int s = 0; s < O.nsp; ++s) {
      if (O.sp[s].child != child || !um_applies(O.sp[s], sd)) continue;
      double a[3]; um_matrix_row(O.sp[s], g, gp, li, weight, a);
      for (int j = 0; j < 3; ++j) if (pos[j] >= 0) val[p0 + pos[j]] += a[j];
    }
  }
}

// The same rows, staged: the values of the 128 consecutive rows of a block are one contiguous span of the value array.  The block
// keeps that span in shared memory, every thread adds its row's element-matrix rows at the precomputed slots, and the span goes
// out (and, when accumulating, comes in) with coalesced accesses.  Spans longer than the image fall back to global updates.
constexpr int UM_JROWS = 128;
constexpr int UM_IMG = UM_JROWS * 20;
__global__ void __launch_bounds__(UM_JROWS) k_um_jacobian_staged(UMDev m, UMSpace S, UMOps O, int child, double weight, int overwrite,
                                                                 const int64_t* __restrict__ rowptr, const uint8_t* __restrict__ slot, double* val)
{
  __shared__ double img[UM_IMG];
  const int64_t d0 = blockIdx.x * (int64_t)UM_JROWS;
  const int nrows = (int)((S.size[child] - d0) < UM_JROWS ? (S.size[child] - d0) : UM_JROWS);
  const int64_t row0 = S.offset[child] + d0;
  const int64_t base = rowptr[row0];
  const int total = (int)(rowptr[row0 + nrows] - base);
  const bool staged = total <= UM_IMG;
  const int tid = threadIdx.x;
  if (staged) {
    if (overwrite) for (int k = tid; k < total; k += UM_JROWS) img[k] = 0.0;
    else for (int k = tid; k < total; k += UM_JROWS) img[k] = val[base + k];
    __syncthreads();
  }
  if (tid < nrows) {
    const int64_t row = row0 + tid;
    const int64_t p0 = rowptr[row];
    double* dst = staged ? img + (p0 - base) : val + p0;
    if (!staged && overwrite) { const int len = (int)(rowptr[row + 1] - p0); for (int k = 0; k < len; ++k) dst[k] = 0.0; }
    const int v = S.d2v[child][d0 + tid];
    for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
      const int cell = m.v2c[e] >> 2, li = m.v2c[e] & 3;
      const uint32_t sd = m.sd[cell];
      bool any = false;
      for (int s = 0; s < O.nsp; ++s) any = any || (O.sp[s].child == child && um_applies(O.sp[s], sd));
      if (!any) continue;
      const Tri g = tri_geometry(m, cell);
      double gp[3][2]; tri_gradients(g, gp);
      const uint8_t s0 = slot[3 * (int64_t)e], s1 = slot[3 * (int64_t)e + 1], s2 = slot[3 * (int64_t)e + 2];
      for (int s = 0; s < O.nsp; ++s) {
        if (O.sp[s].child != child || !um_applies(O.sp[s], sd)) continue;
        double a[3]; um_matrix_row(O.sp[s], g, gp, li, weight, a);
        if (s0 != 255) dst[s0] += a[0];
        if (s1 != 255) dst[s1] += a[1];
        if (s2 != 255) dst[s2] += a[2];
      }
    }
  }
  if (staged) {
    __syncthreads();
    for (int k = tid; k < total; k += UM_JROWS) val[base + k] = img[k];
  }
}

// r[row] += sum over the incident cells of (alpha_volume + lambda_volume)_i
__global__ void k_um_residual_rows(UMDev m, UMSpace S, UMOps O, int child, double weight, double t, const double* __restrict__ x, double* r)
{
  const int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (d >= S.size[child]) return;
  const int v = S.d2v[child][d];
  const int64_t row = S.offset[child] + d;
  double acc = 0.0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const int cell = m.v2c[e] >> 2, li = m.v2c[e] & 3;
    const uint32_t sd = m.sd[cell];
    bool any = false;
    for (int s = 0; s < O.nsp; ++s) any = any || (O.sp[s].child == child && um_applies(O.sp[s], sd));
    if (!any) continue;
    const Tri g = tri_geometry(m, cell);
    double gp[3][2]; tri_gradients(g, gp);
    double xl[3];
    for (int j = 0; j < 3; ++j) xl[j] = x[S.offset[child] + S.v2d[child][m.conn[3 * cell + j]]];
    double rl = 0.0;
    for (int s = 0; s < O.nsp; ++s) {
      const UMSp& sp = O.sp[s];
      if (sp.child != child || !um_applies(sp, sd)) continue;
      int q0; const int nq = tri_rule(sp.quad_order, &q0);
      if (sp.op_id == MDB_OP_POISSON) {               // constant integrand: one term instead of one per quadrature point (see um_matrix_row)
        double gradu[2] = {0.0, 0.0};
        for (int j = 0; j < 3; ++j) for (int dd = 0; dd < 2; ++dd) gradu[dd] += xl[j] * gp[j][dd];
        double dot = 0.0; dot += gradu[0] * pick3(gp[0][0], gp[1][0], gp[2][0], li); dot += gradu[1] * pick3(gp[0][1], gp[1][1], gp[2][1], li);
        rl += weight * (dot * (sp.wsum * g.ie));
      } else {
        for (int q = 0; q < nq; ++q) {
          const double factor = c_tri[q0 + q][2] * g.ie;
          double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
          double u = 0.0; for (int j = 0; j < 3; ++j) u += xl[j] * phi[j];
          rl += weight * (u * pick3(phi[0], phi[1], phi[2], li) * factor);
        }
      }
      if (sp.op_id == MDB_OP_POISSON && sp.P.f.kind != MDB_FN_ZERO)
        for (int q = 0; q < nq; ++q) {
          const double factor = c_tri[q0 + q][2] * g.ie;
          double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
          double xg[2];
          for (int dd = 0; dd < 2; ++dd) xg[dd] = g.p[0][dd] + (g.p[1][dd] - g.p[0][dd]) * c_tri[q0 + q][0] + (g.p[2][dd] - g.p[0][dd]) * c_tri[q0 + q][1];
          const double y = eval_function(sp.P.f, 2, xg, t);
          rl += weight * (-y * pick3(phi[0], phi[1], phi[2], li) * factor);
        }
    }
    acc += rl;
  }
  r[row] += acc;
}

// The same terms cell by cell: one thread per cell evaluates alpha_volume + lambda_volume once (the row kernel evaluates every
// cell once per vertex) and adds its three entries with L2 reductions (RED.ADD.F64) — residual() is additive and the boundary and
// coupling kernels add to the same rows anyway.  A third of the gathers of the row kernel; the summation order inside a row is no
// longer fixed (differences at rounding level, far inside the 1e-12 bar).
__global__ void k_um_residual_cells(UMDev m, UMSpace S, UMOps O, double weight, double t, const double* __restrict__ x, double* r)
{
  const int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (cell >= m.ne) return;
  const uint32_t sd = m.sd[cell];
  bool any = false;
  for (int s = 0; s < O.nsp; ++s) any = any || um_applies(O.sp[s], sd);
  if (!any) return;
  const Tri g = tri_geometry(m, (int)cell);
  double gp[3][2]; tri_gradients(g, gp);
  int vtx[3];
  for (int j = 0; j < 3; ++j) vtx[j] = m.conn[3 * cell + j];
  for (int s = 0; s < O.nsp; ++s) {
    const UMSp& sp = O.sp[s];
    if (!um_applies(sp, sd)) continue;
    int64_t dof[3]; double xl[3];
    bool present = true;
    for (int j = 0; j < 3; ++j) { const int32_t dj = S.v2d[sp.child][vtx[j]]; present = present && dj >= 0; dof[j] = S.offset[sp.child] + dj; }
    if (!present) continue;
    for (int j = 0; j < 3; ++j) xl[j] = x[dof[j]];
    double rl[3] = {0.0, 0.0, 0.0};
    int q0; const int nq = tri_rule(sp.quad_order, &q0);
    if (sp.op_id == MDB_OP_POISSON) {
      double gradu[2] = {0.0, 0.0};
      for (int j = 0; j < 3; ++j) for (int dd = 0; dd < 2; ++dd) gradu[dd] += xl[j] * gp[j][dd];
      const double vol = sp.wsum * g.ie;
      for (int i = 0; i < 3; ++i) { double dot = 0.0; dot += gradu[0] * gp[i][0]; dot += gradu[1] * gp[i][1]; rl[i] += weight * (dot * vol); }
      if (sp.P.f.kind != MDB_FN_ZERO)
        for (int q = 0; q < nq; ++q) {
          const double factor = c_tri[q0 + q][2] * g.ie;
          double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
          double xg[2];
          for (int dd = 0; dd < 2; ++dd) xg[dd] = g.p[0][dd] + (g.p[1][dd] - g.p[0][dd]) * c_tri[q0 + q][0] + (g.p[2][dd] - g.p[0][dd]) * c_tri[q0 + q][1];
          const double y = eval_function(sp.P.f, 2, xg, t);
          for (int i = 0; i < 3; ++i) rl[i] += weight * (-y * phi[i] * factor);
        }
    } else {
      for (int q = 0; q < nq; ++q) {
        const double factor = c_tri[q0 + q][2] * g.ie;
        double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
        double u = 0.0; for (int j = 0; j < 3; ++j) u += xl[j] * phi[j];
        for (int i = 0; i < 3; ++i) rl[i] += weight * (u * phi[i] * factor);
      }
    }
    for (int i = 0; i < 3; ++i) atomicAdd(&r[dof[i]], rl[i]);
  }
}

// lambda_boundary of Poisson on the physical boundary faces (residualassemblerfunctors.hh:123-162).  On a subdomain interface the
// assembler calls the same method with an interior intersection; the predicate answers "no boundary condition" there
// (test/testpoisson.cc:54-57), so only faces without a neighbour do work.
__global__ void k_um_boundary(UMDev m, UMSpace S, UMOps O, const int32_t* __restrict__ bface, int64_t nbface, double weight, double t, double* r)
{
  const int64_t ib = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (ib >= nbface) return;
  const int i = bface[ib];
  const int cell = i / 3, f = i % 3;
  const uint32_t sd = m.sd[cell];
  const int la = c_face_v[f][0], lb = c_face_v[f][1];
  double pa[2], pb[2];
  for (int d = 0; d < 2; ++d) { pa[d] = m.xy[2 * m.conn[3 * cell + la] + d]; pb[d] = m.xy[2 * m.conn[3 * cell + lb] + d]; }
  const double dx = pb[0] - pa[0], dy = pb[1] - pa[1];
  const double ie = sqrt(dx * dx + dy * dy);
  for (int s = 0; s < O.nsp; ++s) {
    const UMSp& sp = O.sp[s];
    if (sp.op_id != MDB_OP_POISSON || !um_applies(sp, sd)) continue;
    double qx[3], qw[3];
    const int nq = line_rule(sp.quad_order, qx, qw);
    double rl[3] = {0.0, 0.0, 0.0};
    for (int q = 0; q < nq; ++q) {
      double xg[2] = {pa[0] + qx[q] * dx, pa[1] + qx[q] * dy};
      if (eval_bctype(sp.P.bc, 2, xg, true) != 1) continue;
      const double lx = c_ref[la][0] + qx[q] * (c_ref[lb][0] - c_ref[la][0]), ly = c_ref[la][1] + qx[q] * (c_ref[lb][1] - c_ref[la][1]);
      double phi[3]; p1_basis(lx, ly, phi);
      const double y = eval_function(sp.P.j, 2, xg, t);
      const double factor = qw[q] * ie;
      for (int k = 0; k < 3; ++k) rl[k] += weight * (y * phi[k] * factor);
    }
    for (int k = 0; k < 3; ++k) {
      const int32_t dd = S.v2d[sp.child][m.conn[3 * cell + k]];
      if (rl[k] != 0.0 && dd >= 0) atomicAdd(&r[S.offset[sp.child] + dd], rl[k]);
    }
  }
}


// ---- cell-wise assembly on per-subproblem cell lists ----------------------------------------------------------------------------
// One thread per (subproblem, cell it applies to).  The lists are built once (um_build_dofmap: vertex ids and container indices of
// the cell's three functions, structure of arrays, cells in Morton order of their centroids so that neighbouring threads touch
// neighbouring rows whatever the order of the file) and, per matrix, the CSR positions of the cell's 9 entries (um_build_pattern:
// "precomputed CSR slots").  A thread reads 12 B of vertex ids + 36 B of slots (coalesced), gathers three coordinate pairs through
// L2, evaluates jacobian_volume once and adds the 9 entries with reductions (RED.ADD.F64 resolves in L2): no adjacency walk, no
// DOF-map gather, no column search.  Summation order inside an entry is not fixed (rounding level, far inside the 1e-12 bar).
struct UMLists { int nlist; int64_t start[9]; UMSp sp[8]; const int32_t* vtx[8]; const int32_t* dof[8]; const int32_t* slot[8]; };

__device__ inline Tri tri_geometry_v(const UMDev& m, int v0, int v1, int v2)
{
  Tri g;
  const double2* xy2 = reinterpret_cast<const double2*>(m.xy);
  const double2 p0 = __ldg(xy2 + v0), p1 = __ldg(xy2 + v1), p2 = __ldg(xy2 + v2);
  g.p[0][0] = p0.x; g.p[0][1] = p0.y; g.p[1][0] = p1.x; g.p[1][1] = p1.y; g.p[2][0] = p2.x; g.p[2][1] = p2.y;
  const double ax = p1.x - p0.x, ay = p1.y - p0.y, bx = p2.x - p0.x, by = p2.y - p0.y;
  const double det = ax * by - ay * bx;
  const double inv = 1.0 / det;
  g.t[0][0] = by * inv; g.t[0][1] = -(ay * inv);
  g.t[1][0] = -(bx * inv); g.t[1][1] = ax * inv;
  g.ie = fabs(det);
  return g;
}
__global__ void __launch_bounds__(256) k_um_jacobian_cells(UMDev m, UMLists L, double weight, double* val)
{
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= L.start[L.nlist]) return;
  int l = 0;
  while (t >= L.start[l + 1]) ++l;
  const int64_t i = t - L.start[l], n = L.start[l + 1] - L.start[l];
  const UMSp& sp = L.sp[l];
  const int32_t* vtx = L.vtx[l]; const int32_t* slot = L.slot[l];
  const Tri g = tri_geometry_v(m, vtx[i], vtx[n + i], vtx[2 * n + i]);
  double gp[3][2]; tri_gradients(g, gp);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double row[3]; um_matrix_row(sp, g, gp, a, weight, row);
#pragma unroll
    for (int b = 0; b < 3; ++b) { const int32_t p = slot[(3 * a + b) * n + i]; if (p >= 0) atomicAdd(val + p, row[b]); }
  }
}
// alpha_volume + lambda_volume, cell by cell on the same lists
__global__ void __launch_bounds__(256) k_um_residual_lists(UMDev m, UMLists L, double weight, double tm, const double* __restrict__ x, double* r)
{
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= L.start[L.nlist]) return;
  int l = 0;
  while (t >= L.start[l + 1]) ++l;
  const int64_t i = t - L.start[l], n = L.start[l + 1] - L.start[l];
  const UMSp& sp = L.sp[l];
  const int32_t* vtx = L.vtx[l]; const int32_t* dofs = L.dof[l];
  const Tri g = tri_geometry_v(m, vtx[i], vtx[n + i], vtx[2 * n + i]);
  double gp[3][2]; tri_gradients(g, gp);
  int32_t dof[3]; double xl[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { dof[j] = dofs[j * n + i]; xl[j] = __ldg(x + dof[j]); }
  double rl[3] = {0.0, 0.0, 0.0};
  int q0; const int nq = tri_rule(sp.quad_order, &q0);
  if (sp.op_id == MDB_OP_POISSON) {
    double gradu[2] = {0.0, 0.0};
    for (int j = 0; j < 3; ++j) for (int dd = 0; dd < 2; ++dd) gradu[dd] += xl[j] * gp[j][dd];
    const double vol = sp.wsum * g.ie;
    for (int a = 0; a < 3; ++a) { double dot = 0.0; dot += gradu[0] * gp[a][0]; dot += gradu[1] * gp[a][1]; rl[a] += weight * (dot * vol); }
    if (sp.P.f.kind != MDB_FN_ZERO)
      for (int q = 0; q < nq; ++q) {
        const double factor = c_tri[q0 + q][2] * g.ie;
        double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
        double xg[2];
        for (int dd = 0; dd < 2; ++dd) xg[dd] = g.p[0][dd] + (g.p[1][dd] - g.p[0][dd]) * c_tri[q0 + q][0] + (g.p[2][dd] - g.p[0][dd]) * c_tri[q0 + q][1];
        const double y = eval_function(sp.P.f, 2, xg, tm);
        for (int a = 0; a < 3; ++a) rl[a] += weight * (-y * phi[a] * factor);
      }
  } else {
    for (int q = 0; q < nq; ++q) {
      const double factor = c_tri[q0 + q][2] * g.ie;
      double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
      double u = 0.0; for (int j = 0; j < 3; ++j) u += xl[j] * phi[j];
      for (int a = 0; a < 3; ++a) rl[a] += weight * (u * phi[a] * factor);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) atomicAdd(r + dof[a], rl[a]);
}
// cell lists: flags -> compaction in Morton order -> vertex ids / container indices (structure of arrays)
__global__ void k_um_list_flags(UMDev m, UMSp sp, const int32_t* __restrict__ order, int32_t* flag)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m.ne) return;
  flag[i] = um_applies(sp, m.sd[order[i]]) ? 1 : 0;
}
__global__ void k_um_list_fill(UMDev m, UMSpace S, UMSp sp, const int32_t* __restrict__ order, const int32_t* __restrict__ flag, const int32_t* __restrict__ scan,
                               int64_t n, int32_t* vtx, int32_t* dof)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m.ne || !flag[i]) return;
  const int cell = order[i]; const int64_t pos = scan[i];
  for (int j = 0; j < 3; ++j) {
    const int v = m.conn[3 * cell + j];
    vtx[j * n + pos] = v;
    dof[j * n + pos] = (int32_t)(S.offset[sp.child] + S.v2d[sp.child][v]);
  }
}
__global__ void k_um_list_slots(int64_t n, const int32_t* __restrict__ dof, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int32_t* slot)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t d[3] = {dof[i], dof[n + i], dof[2 * n + i]};
  for (int a = 0; a < 3; ++a) {
    const int64_t p0 = rowptr[d[a]]; const int len = (int)(rowptr[d[a] + 1] - p0);
    for (int b = 0; b < 3; ++b) { const int pos = um_find(colidx + p0, len, d[b]); slot[(3 * a + b) * n + i] = pos >= 0 ? (int32_t)(p0 + pos) : -1; }
  }
}
// Morton code of the cell centroids (16 bits per axis inside the bounding box)
__global__ void k_um_morton(UMDev m, double x0, double y0, double sx, double sy, uint32_t* code, int32_t* cell)
{
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= m.ne) return;
  double cx = 0.0, cy = 0.0;
  for (int i = 0; i < 3; ++i) { const int v = m.conn[3 * e + i]; cx += m.xy[2 * v]; cy += m.xy[2 * v + 1]; }
  uint32_t ix = (uint32_t)fmin(65535.0, fmax(0.0, (cx / 3.0 - x0) * sx)), iy = (uint32_t)fmin(65535.0, fmax(0.0, (cy / 3.0 - y0) * sy));
  auto spread = [](uint32_t v) { v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu; v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u; return v; };
  code[e] = spread(ix) | (spread(iy) << 1);
  cell[e] = (int32_t)e;
}


// ---- owner-computes rows from per-row pair records -------------------------------------------------------------------------------
// The staged row kernel above walks vertex -> cell -> connectivity -> coordinates per (row, cell) pair: three dependent gathers.
// Here every pair is a 16-byte record built once per (matrix, grid operator): the cell's three vertex ids and a word with the three
// slots inside the row, the row's local index in the cell and the subproblem.  Records are stored warp-transposed (record k of the
// 32 rows of a warp side by side: one coalesced 512-byte load per step, as many steps as the longest row of the warp needs).  A
// thread keeps its row in the block's shared-memory image of the 128 rows' values and the image goes out with coalesced stores: no
// atomics (the reduction unit issues 1.29 cycles per lane, which bounds the cell-wise kernel), no adjacency walk, no zero-fill pass.
struct UMPairPlan { int4* rec = nullptr; int64_t* wptr = nullptr; int64_t nrows = 0, nwarps = 0; };

__global__ void k_um_pair_count(UMDev m, UMSpace S, UMOps O, int64_t ndof, const int32_t* __restrict__ row_child, int32_t* count)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ndof) return;
  const int child = row_child[r];
  const int v = S.d2v[child][r - S.offset[child]];
  int n = 0;
  for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
    const uint32_t sd = m.sd[m.v2c[e] >> 2];
    for (int s = 0; s < O.nsp; ++s) n += (O.sp[s].child == child && um_applies(O.sp[s], sd)) ? 1 : 0;
  }
  count[r] = n;
}
__global__ void k_um_pair_warpmax(int64_t ndof, const int32_t* __restrict__ count, int32_t* wmax)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int n = r < ndof ? count[r] : 0;
  for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
  if ((threadIdx.x & 31) == 0 && (r >> 5) < (ndof + 31) / 32) wmax[r >> 5] = n;
}
__global__ void k_um_pair_fill(UMDev m, UMSpace S, UMOps O, int64_t ndof, const int32_t* __restrict__ row_child, const uint8_t* const* __restrict__ slot_of_child,
                               const int64_t* __restrict__ wptr, int4* rec)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t w = r >> 5;
  if (w >= (ndof + 31) / 32) return;
  const int64_t base = wptr[w]; const int steps = (int)((wptr[w + 1] - base) >> 5);
  int k = 0;
  if (r < ndof) {
    const int child = row_child[r];
    const int v = S.d2v[child][r - S.offset[child]];
    const uint8_t* slot = slot_of_child[child];
    for (int e = m.v2c_ptr[v]; e < m.v2c_ptr[v + 1]; ++e) {
      const int cell = m.v2c[e] >> 2, li = m.v2c[e] & 3;
      const uint32_t sd = m.sd[cell];
      for (int s = 0; s < O.nsp; ++s) {
        if (O.sp[s].child != child || !um_applies(O.sp[s], sd)) continue;
        const uint32_t word = (uint32_t)slot[3 * (int64_t)e] | ((uint32_t)slot[3 * (int64_t)e + 1] << 8) | ((uint32_t)slot[3 * (int64_t)e + 2] << 16) | ((uint32_t)li << 24) | ((uint32_t)s << 26);
        rec[base + (int64_t)k * 32 + lane] = make_int4(m.conn[3 * cell], m.conn[3 * cell + 1], m.conn[3 * cell + 2], (int)word);
        ++k;
      }
    }
  }
  for (; k < steps; ++k) rec[base + (int64_t)k * 32 + lane] = make_int4(-1, -1, -1, 0);
}
// Row i of a cell's element matrix for the pair kernel.  Same quantities as um_matrix_row, arranged for the instruction count (the
// pair kernel is bound by instruction issue, not by HBM): the gradients are never normalised — grad_j . grad_i |det J| =
// (G_j . G_i) / |det J| with G_1 = (b_y, -b_x), G_2 = (-a_y, a_x), G_0 = -(G_1 + G_2) — the reciprocal is rcp.approx refined by two
// Newton steps (1 ulp; the IEEE division is ~25 instructions), products and sums are contracted explicitly (this file is compiled
// with --fmad=false for the bit-matched finite-difference couplings; volume entries are held to 1e-12, not to the bit).
__device__ __forceinline__ double um_rcp(double d)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  r = fma(r, fma(-d, r, 1.0), r);
  return fma(r, fma(-d, r, 1.0), r);
}
__device__ __forceinline__ void um_pair_row(const UMSp& s, const double2 p0, const double2 p1, const double2 p2, int i, double weight, double a[3])
{
  const double ax = p1.x - p0.x, ay = p1.y - p0.y, bx = p2.x - p0.x, by = p2.y - p0.y;
  const double ie = fabs(fma(ax, by, -(ay * bx)));
  if (s.op_id == MDB_OP_POISSON) {
    const double scale = (weight * s.wsum) * um_rcp(ie);
    const double g1x = by, g1y = -bx, g2x = -ay, g2y = ax, g0x = ay - by, g0y = bx - ax;
    const double gix = scale * pick3(g0x, g1x, g2x, i), giy = scale * pick3(g0y, g1y, g2y, i);
    a[0] = fma(g0x, gix, g0y * giy); a[1] = fma(g1x, gix, g1y * giy); a[2] = fma(g2x, gix, g2y * giy);
    return;
  }
  int q0; const int nq = tri_rule(s.quad_order, &q0);
  a[0] = a[1] = a[2] = 0.0;
  for (int q = 0; q < nq; ++q) {
    double phi[3]; p1_basis(c_tri[q0 + q][0], c_tri[q0 + q][1], phi);
    const double f = (weight * (c_tri[q0 + q][2] * ie)) * pick3(phi[0], phi[1], phi[2], i);
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = fma(phi[j], f, a[j]);
  }
}
constexpr int UM_PROWS = 128;
constexpr int UM_PIMG = UM_PROWS * 24;
__global__ void __launch_bounds__(UM_PROWS) k_um_jacobian_pairs(UMDev m, UMOps O, int64_t ndof, const int4* __restrict__ rec, const int64_t* __restrict__ wptr,
                                                                double weight, int overwrite, const int64_t* __restrict__ rowptr, double* val)
{
  __shared__ double img[UM_PIMG];
  const int64_t row0 = blockIdx.x * (int64_t)UM_PROWS;
  const int nrows = (int)((ndof - row0) < UM_PROWS ? (ndof - row0) : UM_PROWS);
  const int64_t vbase = rowptr[row0];
  const int total = (int)(rowptr[row0 + nrows] - vbase);
  const bool staged = total <= UM_PIMG;
  const int tid = threadIdx.x;
  if (staged) {
    if (overwrite) for (int k = tid; k < total; k += UM_PROWS) img[k] = 0.0;
    else for (int k = tid; k < total; k += UM_PROWS) img[k] = val[vbase + k];
    __syncthreads();
  }
  const int64_t w = (row0 + tid) >> 5;
  if (row0 + (tid & ~31) < ndof) {
    const int64_t base = wptr[w]; const int steps = (int)((wptr[w + 1] - base) >> 5);
    const int64_t row = row0 + tid;
    double* dst = nullptr;
    if (tid < nrows) {
      const int64_t p0 = rowptr[row];
      dst = staged ? img + (p0 - vbase) : val + p0;
      if (!staged && overwrite) { const int len = (int)(rowptr[row + 1] - p0); for (int k = 0; k < len; ++k) dst[k] = 0.0; }
    }
    // software pipeline: record k + 2 and the coordinates of pair k + 1 are in flight while pair k is evaluated
    const int lane = tid & 31;
    const double2* xy2 = reinterpret_cast<const double2*>(m.xy);
    const int4 none = make_int4(-1, -1, -1, 0);
    int4 cur = steps > 0 ? __ldg(rec + base + lane) : none;
    int4 nxt = steps > 1 ? __ldg(rec + base + 32 + lane) : none;
    double2 q0 = make_double2(0.0, 0.0), q1 = q0, q2 = q0;
    if (cur.x >= 0) { q0 = __ldg(xy2 + cur.x); q1 = __ldg(xy2 + cur.y); q2 = __ldg(xy2 + cur.z); }
    for (int k = 0; k < steps; ++k) {
      const int4 c4 = cur; const double2 p0 = q0, p1 = q1, p2 = q2;
      cur = nxt;
      nxt = k + 2 < steps ? __ldg(rec + base + (int64_t)(k + 2) * 32 + lane) : none;
      if (cur.x >= 0) { q0 = __ldg(xy2 + cur.x); q1 = __ldg(xy2 + cur.y); q2 = __ldg(xy2 + cur.z); }
      if (c4.x < 0) continue;
      const uint32_t word = (uint32_t)c4.w;
      double a[3]; um_pair_row(O.sp[word >> 26], p0, p1, p2, (int)((word >> 24) & 3u), weight, a);
      const uint32_t s0 = word & 255u, s1 = (word >> 8) & 255u, s2 = (word >> 16) & 255u;
      if (s0 != 255u) dst[s0] += a[0];
      if (s1 != 255u) dst[s1] += a[1];
      if (s2 != 255u) dst[s2] += a[2];
    }
  }
  if (staged) {
    __syncthreads();
    for (int k = tid; k < total; k += UM_PROWS) val[vbase + k] = img[k];
  }
}
__global__ void k_um_row_child(UMSpace S, int64_t ndof, int32_t* row_child)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= ndof) return;
  int k = 0;
  while (k + 1 < S.nchild && r >= S.offset[k + 1]) ++k;
  while (S.size[k] == 0 && k + 1 < S.nchild) ++k;
  row_child[r] = k;
}

// ---- coupling terms ---------------------------------------------------------------------------------------------------------------
struct UMFace {
  double phi1[3][3], phi2[3][3];   // [point][shape function]
  double factor[3];
  double gp1[3][2], gp2[3][2];     // gradients of the shape functions of the inside / outside cell (constant on the cell)
  double normal[2];                // unit outer normal of the inside cell
  double h_F;                      // |corner 0 - corner 1| of the intersection (test/proportionalflowcoupling.hh:135)
  int nq;
};
// face-local quadrature points seen from the inside and the outside cell (geometryInInside / geometryInOutside)
__device__ inline void um_face_setup(const UMDev& m, int cell, int f, int nb, int quad_order, bool gradients, UMFace& F)
{
  const int la = c_face_v[f][0], lb = c_face_v[f][1];
  const int va = m.conn[3 * cell + la], vb = m.conn[3 * cell + lb];
  int ia = 0, ib = 0;
  for (int k = 0; k < 3; ++k) { const int v = m.conn[3 * nb + k]; if (v == va) ia = k; if (v == vb) ib = k; }
  const double dx = m.xy[2 * vb] - m.xy[2 * va], dy = m.xy[2 * vb + 1] - m.xy[2 * va + 1];
  const double ie = sqrt(dx * dx + dy * dy);
  double qx[3], qw[3];
  F.nq = line_rule(quad_order, qx, qw);
  F.h_F = ie;
  if (gradients) {
    const Tri g1 = tri_geometry(m, cell), g2 = tri_geometry(m, nb);
    tri_gradients(g1, F.gp1); tri_gradients(g2, F.gp2);
    // the normal points away from the inside cell's vertex opposite to the face (local vertex 2 - f)
    const int vo = m.conn[3 * cell + (2 - f)];
    double nx = dy / ie, ny = -dx / ie;
    const double side = nx * (m.xy[2 * vo] - m.xy[2 * va]) + ny * (m.xy[2 * vo + 1] - m.xy[2 * va + 1]);
    if (side > 0.0) { nx = -nx; ny = -ny; }
    F.normal[0] = nx; F.normal[1] = ny;
  }
  for (int q = 0; q < F.nq; ++q) {
    const double x1 = c_ref[la][0] + qx[q] * (c_ref[lb][0] - c_ref[la][0]), y1 = c_ref[la][1] + qx[q] * (c_ref[lb][1] - c_ref[la][1]);
    const double x2 = c_ref[ia][0] + qx[q] * (c_ref[ib][0] - c_ref[ia][0]), y2 = c_ref[ia][1] + qx[q] * (c_ref[ib][1] - c_ref[ia][1]);
    p1_basis(x1, y1, F.phi1[q]); p1_basis(x2, y2, F.phi2[q]);
    F.factor[q] = qw[q] * ie;
  }
}
// ProportionalFlowCoupling::alpha_coupling (test/proportionalflowcoupling.hh:27-86) into weighted local views
__device__ inline void um_propflow(const UMFace& F, double k, const double xs[3], const double xn[3], double weight, double rs[3], double rn[3])
{
  for (int q = 0; q < F.nq; ++q) {
    double u_s = 0.0; for (int i = 0; i < 3; ++i) u_s += xs[i] * F.phi1[q][i];
    double u_n = 0.0; for (int i = 0; i < 3; ++i) u_n += xn[i] * F.phi2[q][i];
    const double u_diff = k * (u_s - u_n);
    for (int i = 0; i < 3; ++i) rs[i] += weight * (u_diff * F.phi1[q][i] * F.factor[q]);
    for (int i = 0; i < 3; ++i) rn[i] += weight * (-u_diff * F.phi2[q][i] * F.factor[q]);
  }
}

// ContinuousValueContinuousFlowCoupling::alpha_coupling (test/proportionalflowcoupling.hh:113-233), theta = 1, in the reference's
// operation order (the finite-difference Jacobian differentiates exactly this arithmetic)
__device__ inline void um_cvcf(const UMFace& F, double alpha, const double x1[3], const double x2[3], double weight, double r1[3], double r2[3])
{
  for (int q = 0; q < F.nq; ++q) {
    double gradu1[2] = {0.0, 0.0}, gradu2[2] = {0.0, 0.0};
    for (int i = 0; i < 3; ++i) for (int d = 0; d < 2; ++d) gradu1[d] += x1[i] * F.gp1[i][d];
    for (int i = 0; i < 3; ++i) for (int d = 0; d < 2; ++d) gradu2[d] += x2[i] * F.gp2[i][d];
    double u1 = 0.0; for (int i = 0; i < 3; ++i) u1 += x1[i] * F.phi1[q][i];
    double u2 = 0.0; for (int i = 0; i < 3; ++i) u2 += x2[i] * F.phi2[q][i];
    const double n1[2] = {F.normal[0], F.normal[1]}, n2[2] = {-F.normal[0], -F.normal[1]};
    const double jump_u1 = u1 - u2, jump_u2 = u2 - u1;
    double mean_gradu[2];
    for (int d = 0; d < 2; ++d) { mean_gradu[d] = gradu1[d] + gradu2[d]; mean_gradu[d] *= 0.5; }
    const double factor = F.factor[q];
    for (int i = 0; i < 3; ++i) {
      double mgn = 0.0, gpn = 0.0;
      for (int d = 0; d < 2; ++d) mgn += mean_gradu[d] * n1[d];
      for (int d = 0; d < 2; ++d) gpn += F.gp1[i][d] * n1[d];
      double value = 0.0;
      value -= mgn * F.phi1[q][i];
      value -= 0.5 * gpn * jump_u1;
      value += alpha / F.h_F * jump_u1 * F.phi1[q][i];
      r1[i] += weight * (factor * value);
    }
    for (int i = 0; i < 3; ++i) {
      double mgn = 0.0, gpn = 0.0;
      for (int d = 0; d < 2; ++d) mgn += mean_gradu[d] * n2[d];
      for (int d = 0; d < 2; ++d) gpn += F.gp2[i][d] * n2[d];
      double value = 0.0;
      value -= mgn * F.phi2[q][i];
      value -= 0.5 * gpn * jump_u2;
      value += alpha / F.h_F * jump_u2 * F.phi2[q][i];
      r2[i] += weight * (factor * value);
    }
  }
}
__device__ inline void um_alpha_coupling(const UMCp& cp, const UMFace& F, const double xs[3], const double xn[3], double weight, double rs[3], double rn[3])
{
  if (cp.op_id == MDB_OP_CVCF_COUPLING) um_cvcf(F, cp.intensity, xs, xn, weight, rs, rn);
  else um_propflow(F, cp.intensity, xs, xn, weight, rs, rn);
}

__global__ void k_um_coupling_residual(UMDev m, UMSpace S, UMCp cp, const int32_t* fcell, const int8_t* fface, int64_t nf, double weight,
                                       const double* __restrict__ x, double* r)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nf) return;
  const int cell = fcell[i], f = fface[i], nb = m.nbr[3 * cell + f];
  UMFace F; um_face_setup(m, cell, f, nb, cp.quad_order, cp.op_id == MDB_OP_CVCF_COUPLING, F);
  int64_t gs[3], gn[3]; double xs[3], xn[3];
  for (int k = 0; k < 3; ++k) {
    gs[k] = S.offset[cp.childL] + S.v2d[cp.childL][m.conn[3 * cell + k]]; gn[k] = S.offset[cp.childR] + S.v2d[cp.childR][m.conn[3 * nb + k]];
    xs[k] = x[gs[k]]; xn[k] = x[gn[k]];
  }
  double rs[3] = {0.0, 0.0, 0.0}, rn[3] = {0.0, 0.0, 0.0};
  um_alpha_coupling(cp, F, xs, xn, weight, rs, rn);
  for (int k = 0; k < 3; ++k) { atomicAdd(&r[gs[k]], rs[k]); atomicAdd(&r[gn[k]], rn[k]); }
}

// NumericalJacobianCoupling (couplingutilities.hh:75-135): base line, then one evaluation per coefficient of the inside and of the
// outside cell, entries (up - down) / delta with delta = eps (1 + |u_j|), eps = 1e-8; MDB_JAC_ANALYTIC: columns alpha(e_j) exactly.
__global__ void k_um_coupling_jacobian(UMDev m, UMSpace S, UMCp cp, const int32_t* fcell, const int8_t* fface, int64_t nf, double weight,
                                       const double* __restrict__ x, const int64_t* rowptr, const int32_t* colidx, double* val)
{
  // thread (face, j): the base line and the evaluation with coefficient j perturbed -> column j of the 6 x 6 face block
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= 6 * nf) return;
  const int64_t i = idx / 6; const int j = (int)(idx % 6);
  const int cell = fcell[i], f = fface[i], nb = m.nbr[3 * cell + f];
  UMFace F; um_face_setup(m, cell, f, nb, cp.quad_order, cp.op_id == MDB_OP_CVCF_COUPLING, F);
  int32_t g[6]; double u[6];
  for (int k = 0; k < 3; ++k) {
    g[k] = (int32_t)(S.offset[cp.childL] + S.v2d[cp.childL][m.conn[3 * cell + k]]); g[3 + k] = (int32_t)(S.offset[cp.childR] + S.v2d[cp.childR][m.conn[3 * nb + k]]);
  }
  const bool analytic = cp.jac_mode == MDB_JAC_ANALYTIC;
  for (int k = 0; k < 6; ++k) u[k] = analytic ? 0.0 : x[g[k]];
  double down[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (!analytic) um_alpha_coupling(cp, F, u, u + 3, 1.0, down, down + 3);
  double uj = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) if (k == j) uj = u[k];
  const double delta = analytic ? 1.0 : 1e-8 * (1.0 + fabs(uj));
#pragma unroll
  for (int k = 0; k < 6; ++k) if (k == j) u[k] = analytic ? 1.0 : uj + delta;
  double up[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  um_alpha_coupling(cp, F, u, u + 3, 1.0, up, up + 3);
  int32_t gj = g[0];
#pragma unroll
  for (int k = 1; k < 6; ++k) if (k == j) gj = g[k];
  for (int r = 0; r < 6; ++r) {
    const double entry = analytic ? up[r] : (up[r] - down[r]) / delta;
    const int64_t p0 = rowptr[g[r]];
    const int pos = um_find(colidx + p0, (int)(rowptr[g[r] + 1] - p0), gj);
    if (pos >= 0) atomicAdd(&val[p0 + pos], weight * entry);
  }
}

__global__ void k_um_dirichlet_rows(const int32_t* conlist, int64_t ncon, const int64_t* rowptr, const int32_t* diag, double* val)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ncon) return;
  const int64_t row = conlist[i];
  const int64_t p0 = rowptr[row], p1 = rowptr[row + 1];
  for (int64_t p = p0; p < p1; ++p) val[p] = 0.0;
  if (diag[row] >= 0) val[p0 + diag[row]] = 1.0;
}

int64_t um_scan(Ctx* c, const int32_t* d_in, int32_t* d_out, int64_t n)
{
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

template <typename T> void um_dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

} // anonymous namespace

// ---- host entry points ----------------------------------------------------------------------------------------------------------------
void um_free(Ctx* c)
{
  if (!c->um) return;
  UMesh* u = c->um;
  ufe_free(c);
  um_dfree(u->d_xy); um_dfree(u->d_conn); um_dfree(u->d_nbr); um_dfree(u->d_v2c_ptr); um_dfree(u->d_v2c); um_dfree(u->d_bface);
  um_dfree(u->d_cell_edge); um_dfree(u->d_group); um_dfree(u->d_cell_order);
  delete u; c->um = nullptr;
}

// Mesh ingest: validation and the connectivity a grid manager provides (face neighbours, edge numbers, vertex -> cell), built once
// ON THE DEVICE from the caller's coordinate / connectivity arrays (round 1 and the first sessions of round 2 sorted 3 ne edge keys
// with std::sort on the host: 1.5 - 3 s at 8.4 M triangles).  Everything is a sort or a scan:
//   * face neighbours and edge numbers: 64-bit keys (min vertex << 32 | max vertex) of the 3 ne (cell, face) positions, stable radix
//     sort with the position as payload -> the members of a group are adjacent and ascending in position; two members are each
//     other's neighbour, three are a non-manifold mesh; the first member of a group is the edge's first appearance in (cell, local
//     face) order, a scan over the first-appearance flags numbers the edges in that order (the P2 numbering of ufe.cu);
//   * vertex -> cell: keys = vertex, payload = 4 cell + local vertex in position order, stable sort -> cells ascending per vertex;
//     the offsets are a scan over integer counts;
//   * boundary faces: positions without neighbour, compacted in ascending order.
// The host keeps nothing: it only reads back the error word and three counts.
enum { UM_ERR_RANGE = 1, UM_ERR_DEGENERATE = 2, UM_ERR_MANIFOLD = 4 };
__global__ void k_um_ingest_cells(int64_t nv, int64_t ne, const int64_t* __restrict__ conn64, const double* __restrict__ xy, int32_t* conn, uint64_t* key,
                                  int32_t* pos, uint32_t* vkey, int32_t* vpay, int32_t* vcount, int* err)
{
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int64_t v[3]; bool ok = true;
  for (int k = 0; k < 3; ++k) { v[k] = conn64[3 * e + k]; ok = ok && v[k] >= 0 && v[k] < nv; }
  if (!ok) {                                       // keep the arrays well defined; the call fails on the error word
    atomicOr(err, UM_ERR_RANGE);
    for (int k = 0; k < 3; ++k) { conn[3 * e + k] = 0; key[3 * e + k] = ~0ull; pos[3 * e + k] = (int32_t)(3 * e + k); vkey[3 * e + k] = 0u; vpay[3 * e + k] = (int32_t)(4 * e + k); }
    return;
  }
  const double ax = xy[2 * v[1]] - xy[2 * v[0]], ay = xy[2 * v[1] + 1] - xy[2 * v[0] + 1];
  const double bx = xy[2 * v[2]] - xy[2 * v[0]], by = xy[2 * v[2] + 1] - xy[2 * v[0] + 1];
  if (v[0] == v[1] || v[0] == v[2] || v[1] == v[2] || ax * by - ay * bx == 0.0) atomicOr(err, UM_ERR_DEGENERATE);
  for (int k = 0; k < 3; ++k) {
    conn[3 * e + k] = (int32_t)v[k];
    vkey[3 * e + k] = (uint32_t)v[k]; vpay[3 * e + k] = (int32_t)(4 * e + k);
    atomicAdd(vcount + v[k], 1);
  }
  for (int f = 0; f < 3; ++f) {
    uint64_t a = (uint64_t)v[c_face_v[f][0]], b = (uint64_t)v[c_face_v[f][1]];
    if (a > b) { const uint64_t t = a; a = b; b = t; }
    key[3 * e + f] = (a << 32) | b; pos[3 * e + f] = (int32_t)(3 * e + f);
  }
}
// sorted edge groups -> neighbours, first-appearance flags, the position of every member's first appearance
__global__ void k_um_ingest_groups(int64_t n3, const uint64_t* __restrict__ key, const int32_t* __restrict__ pos, int32_t* nbr, int32_t* first_flag,
                                   int32_t* first_pos, int* err)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const uint64_t k = key[i];
  const bool head = i == 0 || key[i - 1] != k;
  const bool more = i + 1 < n3 && key[i + 1] == k;
  const int32_t p = pos[i];
  if (head) {
    if (more && i + 2 < n3 && key[i + 2] == k) atomicOr(err, UM_ERR_MANIFOLD);
    nbr[p] = more ? pos[i + 1] / 3 : -1;
    first_flag[p] = 1; first_pos[p] = p;
  } else {
    nbr[p] = pos[i - 1] / 3;
    first_flag[p] = 0; first_pos[p] = pos[i - 1];      // valid for groups of two (anything longer is an error)
  }
}
__global__ void k_um_ingest_edges(int64_t n3, const int32_t* __restrict__ first_pos, const int32_t* __restrict__ first_scan, const int32_t* __restrict__ nbr,
                                  int32_t* cell_edge, int32_t* bflag)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n3) return;
  cell_edge[i] = first_scan[first_pos[i]];
  bflag[i] = nbr[i] < 0 ? 1 : 0;
}
__global__ void k_um_ingest_bfaces(int64_t n3, const int32_t* __restrict__ bflag, const int32_t* __restrict__ bscan, int32_t* bface)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n3 && bflag[i]) bface[bscan[i]] = (int32_t)i;
}
__global__ void k_um_bbox(int64_t nv, const double* __restrict__ xy, double* box)     // box = {min x, min y, -max x, -max y} kept as minima
{
  double lo0 = 1e300, lo1 = 1e300, hi0 = -1e300, hi1 = -1e300;
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < nv; v += (int64_t)gridDim.x * blockDim.x) {
    const double x = xy[2 * v], y = xy[2 * v + 1];
    lo0 = fmin(lo0, x); lo1 = fmin(lo1, y); hi0 = fmax(hi0, x); hi1 = fmax(hi1, y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo0 = fmin(lo0, __shfl_xor_sync(0xffffffffu, lo0, o)); lo1 = fmin(lo1, __shfl_xor_sync(0xffffffffu, lo1, o));
    hi0 = fmax(hi0, __shfl_xor_sync(0xffffffffu, hi0, o)); hi1 = fmax(hi1, __shfl_xor_sync(0xffffffffu, hi1, o));
  }
  if ((threadIdx.x & 31) == 0) {
    // order-preserving integer image of a double: min / max of the images = image of the min / max (exact, order independent)
    auto enc = [](double d) { const long long b = __double_as_longlong(d); return b < 0 ? b ^ 0x7fffffffffffffffll : b; };
    atomicMin((long long*)box + 0, enc(lo0)); atomicMin((long long*)box + 1, enc(lo1));
    atomicMax((long long*)box + 2, enc(hi0)); atomicMax((long long*)box + 3, enc(hi1));
  }
}

void um_ingest(Ctx* c, int64_t nv, const double* xy, int64_t ne, const int64_t* conn)
{
  MDB_REQUIRE(nv >= 3 && ne >= 1 && nv < (int64_t)1 << 30 && ne < (int64_t)1 << 29, MDB_EINVAL, "mdb_mesh_unstructured: bad vertex / cell count");
  const int64_t n3 = 3 * ne;
  // inputs and the arrays that stay
  double* d_xy = dalloc<double>(2 * nv); int32_t* d_conn = dalloc<int32_t>(n3); int32_t* d_nbr = dalloc<int32_t>(n3);
  int32_t* d_v2c_ptr = dalloc<int32_t>(nv + 1); int32_t* d_v2c = dalloc<int32_t>(n3); int32_t* d_cell_edge = dalloc<int32_t>(n3);
  // scratch
  int64_t* d_conn64 = dalloc<int64_t>(n3);
  uint64_t* d_key = dalloc<uint64_t>(n3); uint64_t* d_key2 = dalloc<uint64_t>(n3);
  int32_t* d_pos = dalloc<int32_t>(n3); int32_t* d_pos2 = dalloc<int32_t>(n3);
  uint32_t* d_vkey = dalloc<uint32_t>(n3); uint32_t* d_vkey2 = dalloc<uint32_t>(n3); int32_t* d_vpay = dalloc<int32_t>(n3);
  int32_t* d_vcount = dalloc<int32_t>(nv + 1);
  int32_t* d_flag = dalloc<int32_t>(n3); int32_t* d_scan = dalloc<int32_t>(n3); int32_t* d_fpos = dalloc<int32_t>(n3);
  int* d_err = dalloc<int>(1);
  int32_t* d_bface = nullptr;
  auto release = [&](bool all) {
    cudaFree(d_conn64); cudaFree(d_key); cudaFree(d_key2); cudaFree(d_pos); cudaFree(d_pos2); cudaFree(d_vkey); cudaFree(d_vkey2); cudaFree(d_vpay);
    cudaFree(d_vcount); cudaFree(d_flag); cudaFree(d_scan); cudaFree(d_fpos); cudaFree(d_err);
    if (all) { cudaFree(d_xy); cudaFree(d_conn); cudaFree(d_nbr); cudaFree(d_v2c_ptr); cudaFree(d_v2c); cudaFree(d_cell_edge); if (d_bface) cudaFree(d_bface); }
  };
  int64_t nedges = 0, nbface = 0;
  try {
    MDB_CUDA(cudaMemcpyAsync(d_xy, xy, 2 * nv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    MDB_CUDA(cudaMemcpyAsync(d_conn64, conn, n3 * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    MDB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), c->stream));
    MDB_CUDA(cudaMemsetAsync(d_vcount, 0, (nv + 1) * sizeof(int32_t), c->stream));
    MDB_LAUNCH(c, k_um_ingest_cells, cdiv(ne, 256), 256, 0, nv, ne, d_conn64, d_xy, d_conn, d_key, d_pos, d_vkey, d_vpay, d_vcount, d_err);
    int err = 0;
    MDB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    MDB_REQUIRE(!(err & UM_ERR_RANGE), MDB_EINVAL, "mdb_mesh_unstructured: vertex index out of range");
    MDB_REQUIRE(!(err & UM_ERR_DEGENERATE), MDB_EINVAL, "mdb_mesh_unstructured: degenerate cell");
    // the two sorts share one temporary buffer
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    const int vbits = 32 - __builtin_clz((unsigned)std::max<int64_t>(nv - 1, 1));
    const int kbits = 32 + vbits;
    cub::DeviceRadixSort::SortPairs(nullptr, tb1, d_key, d_key2, d_pos, d_pos2, (int)n3, 0, kbits, c->stream);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, d_vkey, d_vkey2, d_vpay, d_v2c, (int)n3, 0, vbits, c->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_vcount, d_v2c_ptr, (int)(nv + 1), c->stream);
    const size_t tb = std::max(std::max(tb1, tb2), std::max(tb3, (size_t)1));
    void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb));
    cudaError_t e1 = cub::DeviceRadixSort::SortPairs(d_t, tb1, d_key, d_key2, d_pos, d_pos2, (int)n3, 0, kbits, c->stream); c->launches++;
    cudaError_t e2 = cub::DeviceRadixSort::SortPairs(d_t, tb2, d_vkey, d_vkey2, d_vpay, d_v2c, (int)n3, 0, vbits, c->stream); c->launches++;   // stable: cells ascending per vertex
    cudaError_t e3 = cub::DeviceScan::ExclusiveSum(d_t, tb3, d_vcount, d_v2c_ptr, (int)(nv + 1), c->stream); c->launches++;
    cudaError_t e4 = cudaStreamSynchronize(c->stream);
    cudaFree(d_t);
    MDB_CUDA(e1); MDB_CUDA(e2); MDB_CUDA(e3); MDB_CUDA(e4);
    MDB_LAUNCH(c, k_um_ingest_groups, cdiv(n3, 256), 256, 0, n3, d_key2, d_pos2, d_nbr, d_flag, d_fpos, d_err);
    nedges = um_scan(c, d_flag, d_scan, n3);
    MDB_CUDA(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    MDB_REQUIRE(!(err & UM_ERR_MANIFOLD), MDB_EINVAL, "mdb_mesh_unstructured: more than two cells share a face (non-manifold mesh)");
    MDB_LAUNCH(c, k_um_ingest_edges, cdiv(n3, 256), 256, 0, n3, d_fpos, d_scan, d_nbr, d_cell_edge, d_flag);       // d_flag becomes the boundary flag
    nbface = um_scan(c, d_flag, d_pos, n3);                                                                        // d_pos is free after the sort
    d_bface = dalloc<int32_t>(nbface);
    if (nbface > 0) MDB_LAUNCH(c, k_um_ingest_bfaces, cdiv(n3, 256), 256, 0, n3, d_flag, d_pos, d_bface);
    MDB_CUDA(cudaStreamSynchronize(c->stream));
  } catch (...) { release(true); throw; }
  release(false);
  um_free(c);
  UMesh* u = new UMesh(); c->um = u;
  u->nv = nv; u->ne = ne;
  u->d_xy = d_xy; u->d_conn = d_conn; u->d_nbr = d_nbr; u->d_v2c_ptr = d_v2c_ptr; u->d_v2c = d_v2c;
  u->nedges = nedges; u->d_cell_edge = d_cell_edge; u->d_group = dalloc<int32_t>(ne);
  u->nbface = nbface; u->d_bface = d_bface;
  MDB_CUDA(cudaMemsetAsync(u->d_group, 0, ne * 4, c->stream));
  c->h2d_bytes += 2 * nv * 8 + n3 * 8;
  // processing order of the cell-wise kernels: Morton order of the centroids (the numbering itself is the caller's and stays)
  {
    // bounding box on the device (exact and order independent: integer minima / maxima of an order-preserving image)
    auto enc = [](double d) { long long b; std::memcpy(&b, &d, 8); return b < 0 ? b ^ 0x7fffffffffffffffll : b; };
    auto dec = [](long long b) { if (b < 0) b ^= 0x7fffffffffffffffll; double d; std::memcpy(&d, &b, 8); return d; };
    long long hbox[4] = {enc(1e300), enc(1e300), enc(-1e300), enc(-1e300)};
    double* d_box = dalloc<double>(4);
    MDB_CUDA(cudaMemcpyAsync(d_box, hbox, sizeof(hbox), cudaMemcpyHostToDevice, c->stream));
    MDB_LAUNCH(c, k_um_bbox, std::min<int64_t>(cdiv(nv, 256), 148 * 8), 256, 0, nv, u->d_xy, d_box);
    MDB_CUDA(cudaMemcpyAsync(hbox, d_box, sizeof(hbox), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_box);
    const double lo[2] = {dec(hbox[0]), dec(hbox[1])}, hi[2] = {dec(hbox[2]), dec(hbox[3])};
    uint32_t* d_code = dalloc<uint32_t>(ne); uint32_t* d_code2 = dalloc<uint32_t>(ne); int32_t* d_cell = dalloc<int32_t>(ne);
    u->d_cell_order = dalloc<int32_t>(ne);
    UMDev m = um_dev(c);
    MDB_LAUNCH(c, k_um_morton, cdiv(ne, 256), 256, 0, m, lo[0], lo[1], 65535.0 / std::max(hi[0] - lo[0], 1e-300), 65535.0 / std::max(hi[1] - lo[1], 1e-300), d_code, d_cell);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, d_code, d_code2, d_cell, u->d_cell_order, (int)ne, 0, 32, c->stream);
    void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
    MDB_CUDA(cub::DeviceRadixSort::SortPairs(d_t, tb, d_code, d_code2, d_cell, u->d_cell_order, (int)ne, 0, 32, c->stream)); c->launches++;
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t); cudaFree(d_code); cudaFree(d_code2); cudaFree(d_cell);
  }
}

void um_build_dofmap(Ctx* c)
{
  const UMDev m = um_dev(c);
  int32_t* d_flag = dalloc<int32_t>(m.nv); int32_t* d_scan = dalloc<int32_t>(m.nv);
  int64_t offset = 0;
  for (Child& ch : c->children) {
    MDB_REQUIRE(ch.fe_kind == MDB_FE_P1, MDB_EINVAL, "unstructured mesh: only MDB_FE_P1 child spaces");
    ch.nlat = m.nv; ch.lat_n[0] = (int)m.nv; ch.lat_n[1] = ch.lat_n[2] = 1;
    um_dfree(ch.d_lat2dof); um_dfree(ch.d_dof2lat);
    ch.d_lat2dof = dalloc<int32_t>(m.nv);
    MDB_LAUNCH(c, k_um_mark, cdiv(m.nv, 256), 256, 0, m, ch.subdomain, d_flag);
    ch.size = um_scan(c, d_flag, d_scan, m.nv);
    ch.d_dof2lat = dalloc<int32_t>(ch.size);
    MDB_LAUNCH(c, k_um_assign, cdiv(m.nv, 256), 256, 0, m.nv, d_flag, d_scan, ch.d_lat2dof, ch.d_dof2lat);
    ch.offset = offset; offset += ch.size;
  }
  MDB_REQUIRE(offset < (int64_t)2000000000, MDB_EINVAL, "unstructured mesh: more than 2e9 DOFs per context");
  c->ndof = offset;
  um_dfree(c->d_constrained);
  c->d_constrained = dalloc<uint8_t>(c->ndof);
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_flag); cudaFree(d_scan);
  // cell lists of the subproblems (cell-wise kernels)
  {
    const UMSpace S = um_space(c);
    int32_t* d_f = dalloc<int32_t>(m.ne); int32_t* d_s = dalloc<int32_t>(m.ne);
    for (size_t si = 0; si < c->sps.size(); ++si) {
      SubProblemDesc& d = c->sps[si];
      const UMOps O = um_ops(c, std::vector<int>(1, (int)si), std::vector<int>());
      um_dfree(d.d_um_vtx); um_dfree(d.d_um_dof);
      MDB_LAUNCH(c, k_um_list_flags, cdiv(m.ne, 256), 256, 0, m, O.sp[0], c->um->d_cell_order, d_f);
      d.um_ncells = um_scan(c, d_f, d_s, m.ne);
      d.d_um_vtx = dalloc<int32_t>(3 * d.um_ncells); d.d_um_dof = dalloc<int32_t>(3 * d.um_ncells);
      if (d.um_ncells > 0) MDB_LAUNCH(c, k_um_list_fill, cdiv(m.ne, 256), 256, 0, m, S, O.sp[0], c->um->d_cell_order, d_f, d_s, d.um_ncells, d.d_um_vtx, d.d_um_dof);
    }
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_f); cudaFree(d_s);
  }
}

// parity export: the MultiDomainLocalFunctionSpace of every cell, children concatenated (multidomainlocalfunctionspace.hh:348-358)
void um_export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs)
{
  const UMesh* u = c->um;
  std::vector<int32_t> conn(3 * u->ne); std::vector<uint8_t> sd(u->ne);
  MDB_CUDA(cudaMemcpy(conn.data(), u->d_conn, 3 * u->ne * 4, cudaMemcpyDeviceToHost));
  MDB_CUDA(cudaMemcpy(sd.data(), c->d_cell_sd, u->ne, cudaMemcpyDeviceToHost));
  std::vector<std::vector<int32_t>> v2d(c->children.size());
  for (size_t k = 0; k < c->children.size(); ++k) {
    v2d[k].resize(u->nv);
    MDB_CUDA(cudaMemcpy(v2d[k].data(), c->children[k].d_lat2dof, u->nv * 4, cudaMemcpyDeviceToHost));
  }
  int64_t p = 0;
  for (int64_t e = 0; e < u->ne; ++e) {
    cell_ptr[e] = p;
    for (size_t k = 0; k < c->children.size(); ++k) {
      const Child& ch = c->children[k];
      if (!(ch.subdomain < 0 || ((sd[e] >> ch.subdomain) & 1))) continue;
      for (int i = 0; i < 3; ++i) { if (cell_dofs) cell_dofs[p] = ch.offset + v2d[k][conn[3 * e + i]]; ++p; }
    }
  }
  cell_ptr[u->ne] = p;
}

static UMBcs um_bcs(Ctx* c, const int* sps, const mdb_bctype* bcs, const mdb_function* fns, int n)
{
  MDB_REQUIRE(n <= 8, MDB_EINVAL, "unstructured mesh: at most 8 subproblems per call");
  UMBcs B; B.n = n;
  for (int i = 0; i < n; ++i) {
    const SubProblemDesc& d = c->sps[sps[i]];
    B.sp[i].op_id = d.op_id; B.sp[i].child = d.children[0]; B.sp[i].cond_kind = d.cond_kind; B.sp[i].mask = d.mask; B.sp[i].quad_order = 0; B.sp[i].wsum = 0.5; B.sp[i].P = d.poisson;
    if (bcs) B.bc[i] = bcs[i]; else B.bc[i] = mdb_bctype{};
    if (fns) B.fn[i] = fns[i]; else B.fn[i] = mdb_function{};
  }
  return B;
}

void um_assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n)
{
  const UMDev m = um_dev(c);
  MDB_CUDA(cudaMemsetAsync(c->d_constrained, 0, c->ndof, c->stream));
  MDB_LAUNCH(c, k_um_constraints, cdiv(3 * m.ne, 256), 256, 0, m, um_space(c), um_bcs(c, sps, bcs, nullptr, n), c->d_constrained);
}

void um_interpolate(Ctx* c, const int* sps, const mdb_function* fns, int n, double t, double* d_x)
{
  const UMDev m = um_dev(c);
  const UMSpace S = um_space(c);
  const UMBcs B = um_bcs(c, sps, nullptr, fns, n);
  for (int k = 0; k < S.nchild; ++k)
    if (S.size[k] > 0) MDB_LAUNCH(c, k_um_interpolate, cdiv(S.size[k], 256), 256, 0, m, S, B, k, t, d_x);
}

void um_build_facelists(Ctx* c, GridOp& go)
{
  const UMDev m = um_dev(c);
  const UMOps O = um_ops(c, go.sps, go.cps);
  const int64_t n = 3 * m.ne;
  int32_t* d_flag = dalloc<int32_t>(n); int32_t* d_scan = dalloc<int32_t>(n);
  go.faces.clear();
  for (int q = 0; q < O.ncp; ++q) {
    MDB_LAUNCH(c, k_um_face_flags, cdiv(n, 256), 256, 0, m, O.cp[q], d_flag);
    FaceList fl; fl.n = um_scan(c, d_flag, d_scan, n);
    fl.d_cell = dalloc<int32_t>(fl.n); fl.d_face = dalloc<int8_t>(fl.n);
    MDB_LAUNCH(c, k_um_face_compact, cdiv(n, 256), 256, 0, n, d_flag, d_scan, fl.d_cell, fl.d_face);
    go.faces.push_back(fl);
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_flag); cudaFree(d_scan);
}

void um_build_pattern(Ctx* c, Matrix& M)
{
  const UMDev m = um_dev(c);
  const UMSpace S = um_space(c);
  std::vector<int> sps, cps;
  for (int g : M.gos) { for (int s : c->gos[g].sps) if (std::find(sps.begin(), sps.end(), s) == sps.end()) sps.push_back(s);
                        for (int q : c->gos[g].cps) if (std::find(cps.begin(), cps.end(), q) == cps.end()) cps.push_back(q); }
  const UMOps O = um_ops(c, sps, cps);
  const int64_t n = c->ndof;
  int32_t* d_count = dalloc<int32_t>(n); int* d_over = dalloc<int>(1);
  MDB_CUDA(cudaMemsetAsync(d_count, 0, n * 4, c->stream));
  MDB_CUDA(cudaMemsetAsync(d_over, 0, 4, c->stream));
  for (int k = 0; k < S.nchild; ++k)
    if (S.size[k] > 0) MDB_LAUNCH(c, k_um_pattern_count, cdiv(S.size[k], 128), 128, 0, m, S, O, k, d_count, d_over);
  int over = 0;
  MDB_CUDA(cudaMemcpyAsync(&over, d_over, 4, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_over);
  if (over) { cudaFree(d_count); throw Error(MDB_EINVAL, "unstructured mesh: a matrix row has more than 128 entries (vertex valence too high)"); }
  // row pointers: 64-bit scan of the counts
  int64_t* d_c64 = dalloc<int64_t>(n); int64_t* d_s64 = dalloc<int64_t>(n);
  MDB_LAUNCH(c, k_um_to64, cdiv(n, 256), 256, 0, n, d_count, d_c64);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_c64, d_s64, (int)n, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_c64, d_s64, (int)n, c->stream)); c->launches++;
  M.d_rowptr = dalloc<int64_t>(n + 1);
  MDB_LAUNCH(c, k_um_widen, cdiv(n, 256), 256, 0, n, d_count, d_s64, M.d_rowptr);
  int64_t nnz = 0;
  MDB_CUDA(cudaMemcpyAsync(&nnz, M.d_rowptr + n, 8, cudaMemcpyDeviceToHost, c->stream));
  int32_t maxrow = 0;
  {
    size_t tb2 = 0; int32_t* d_max = dalloc<int32_t>(1);
    cub::DeviceReduce::Max(nullptr, tb2, d_count, d_max, (int)n, c->stream);
    void* d_t2 = nullptr; MDB_CUDA(cudaMalloc(&d_t2, tb2 ? tb2 : 1));
    MDB_CUDA(cub::DeviceReduce::Max(d_t2, tb2, d_count, d_max, (int)n, c->stream)); c->launches++;
    MDB_CUDA(cudaMemcpyAsync(&maxrow, d_max, 4, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_t2); cudaFree(d_max);
  }
  cudaFree(d_t); cudaFree(d_c64); cudaFree(d_s64); cudaFree(d_count);
  M.nnz = nnz; M.max_row = maxrow;
  M.d_colidx = dalloc<int32_t>(nnz); M.d_val = dalloc<double>(nnz); M.d_diag = dalloc<int32_t>(n);
  MDB_CUDA(cudaMemsetAsync(M.d_val, 0, nnz * sizeof(double), c->stream));
  for (int k = 0; k < S.nchild; ++k) {
    M.child_begin[k] = S.offset[k]; M.child_end[k] = S.offset[k] + S.size[k];
    if (S.size[k] == 0) continue;
    uint8_t* d_slot = dalloc<uint8_t>(9 * m.ne);
    MDB_CUDA(cudaMemsetAsync(d_slot, 255, 9 * m.ne, c->stream));
    M.slot_tables[std::make_pair(k, -1)] = d_slot;            // freed with the matrix
    MDB_LAUNCH(c, k_um_pattern_fill, cdiv(S.size[k], 128), 128, 0, m, S, O, k, M.d_rowptr, M.d_colidx, M.d_diag, d_slot);
  }
  // precomputed CSR slots of the cell-wise Jacobian kernel: 9 absolute positions per (subproblem, cell)
  if (nnz < ((int64_t)1 << 31))
    for (int s_id : sps) {
      const SubProblemDesc& d = c->sps[s_id];
      int32_t* d_slots = dalloc<int32_t>(9 * d.um_ncells);
      M.slot_tables[std::make_pair(s_id, -2)] = reinterpret_cast<uint8_t*>(d_slots);
      if (d.um_ncells > 0) MDB_LAUNCH(c, k_um_list_slots, cdiv(d.um_ncells, 128), 128, 0, d.um_ncells, d.d_um_dof, M.d_rowptr, M.d_colidx, d_slots);
    }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
}

void um_pair_plan_free(void* p)
{
  UMPairPlan* P = (UMPairPlan*)p;
  if (!P) return;
  if (P->rec) cudaFree(P->rec);
  if (P->wptr) cudaFree(P->wptr);
  delete P;
}
__global__ void k_um_times32(int64_t n, const int32_t* in, int64_t* out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = 32 * (int64_t)in[i];
}
__global__ void k_um_close_scan(int64_t n, const int64_t* sizes, int64_t* scan)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) scan[n] = scan[n - 1] + sizes[n - 1];
}
static UMPairPlan* um_pair_plan(Ctx* c, const GridOp& go, Matrix& M)
{
  std::vector<int> key(go.sps);
  auto it = M.um_pair_plans.find(key);
  if (it != M.um_pair_plans.end()) return (UMPairPlan*)it->second;
  const UMDev m = um_dev(c);
  const UMSpace S = um_space(c);
  const UMOps O = um_ops(c, go.sps, std::vector<int>());
  if (O.nsp > 7 || M.max_row > 255) { M.um_pair_plans[key] = nullptr; return nullptr; }
  // every child of the operator needs its slot table (written by k_um_pattern_fill)
  std::vector<const uint8_t*> h_slot(MDB_MAX_CHILDREN, nullptr);
  for (int k = 0; k < S.nchild; ++k) { auto st = M.slot_tables.find(std::make_pair(k, -1)); if (st != M.slot_tables.end()) h_slot[k] = st->second; }
  for (int s = 0; s < O.nsp; ++s) if (!h_slot[O.sp[s].child]) { M.um_pair_plans[key] = nullptr; return nullptr; }
  const int64_t n = c->ndof, nw = (n + 31) / 32;
  UMPairPlan* P = new UMPairPlan(); P->nrows = n; P->nwarps = nw;
  int32_t* d_child = dalloc<int32_t>(n); int32_t* d_count = dalloc<int32_t>(n); int32_t* d_wmax = dalloc<int32_t>(nw);
  int64_t* d_size = dalloc<int64_t>(nw);
  const uint8_t** d_slot = nullptr; MDB_CUDA(cudaMalloc(&d_slot, MDB_MAX_CHILDREN * sizeof(uint8_t*)));
  MDB_CUDA(cudaMemcpyAsync(d_slot, h_slot.data(), MDB_MAX_CHILDREN * sizeof(uint8_t*), cudaMemcpyHostToDevice, c->stream));
  MDB_LAUNCH(c, k_um_row_child, cdiv(n, 256), 256, 0, S, n, d_child);
  MDB_LAUNCH(c, k_um_pair_count, cdiv(n, 256), 256, 0, m, S, O, n, d_child, d_count);
  MDB_LAUNCH(c, k_um_pair_warpmax, cdiv(nw * 32, 256), 256, 0, n, d_count, d_wmax);
  MDB_LAUNCH(c, k_um_times32, cdiv(nw, 256), 256, 0, nw, d_wmax, d_size);
  P->wptr = dalloc<int64_t>(nw + 1);
  size_t tb = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, d_size, P->wptr, (int)nw, c->stream);
  void* d_t = nullptr; MDB_CUDA(cudaMalloc(&d_t, tb ? tb : 1));
  MDB_CUDA(cub::DeviceScan::ExclusiveSum(d_t, tb, d_size, P->wptr, (int)nw, c->stream)); c->launches++;
  MDB_LAUNCH(c, k_um_close_scan, 1, 32, 0, nw, d_size, P->wptr);
  int64_t nrec = 0;
  MDB_CUDA(cudaMemcpyAsync(&nrec, P->wptr + nw, 8, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  P->rec = dalloc<int4>(nrec);
  MDB_LAUNCH(c, k_um_pair_fill, cdiv(nw * 32, 256), 256, 0, m, S, O, n, d_child, d_slot, P->wptr, P->rec);
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_t); cudaFree(d_child); cudaFree(d_count); cudaFree(d_wmax); cudaFree(d_size); cudaFree(d_slot);
  M.um_pair_plans[key] = P;
  return P;
}

static bool um_lists(Ctx* c, const GridOp& go, const Matrix* M, UMLists& L)
{
  if (go.sps.empty() || go.sps.size() > 8) return false;
  const UMOps O = um_ops(c, go.sps, std::vector<int>());
  L.nlist = (int)go.sps.size(); L.start[0] = 0;
  for (int i = 0; i < L.nlist; ++i) {
    const SubProblemDesc& d = c->sps[go.sps[i]];
    if (!d.d_um_vtx) return false;
    L.sp[i] = O.sp[i]; L.vtx[i] = d.d_um_vtx; L.dof[i] = d.d_um_dof; L.slot[i] = nullptr;
    L.start[i + 1] = L.start[i] + d.um_ncells;
    if (M) {
      auto it = M->slot_tables.find(std::make_pair(go.sps[i], -2));
      if (it == M->slot_tables.end()) return false;
      L.slot[i] = reinterpret_cast<const int32_t*>(it->second);
    }
  }
  return true;
}

void um_residual(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r)
{
  const UMDev m = um_dev(c);
  const UMSpace S = um_space(c);
  const UMOps O = um_ops(c, go.sps, go.cps);
  static const bool by_rows = getenv("MDB_UM_RES_ROWS") != nullptr;       // A/B switch: the owner-computes row kernel (fixed summation order)
  static const bool by_cells = getenv("MDB_UM_RES_CELLS") != nullptr;     // A/B switch: the cell kernel that walks the mesh arrays (no lists)
  static const bool serial = getenv("MDB_UM_RES_SERIAL") != nullptr;      // A/B switch: boundary and coupling faces after the cells
  auto faces = [&]() {
    c->phase_begin("residual_boundary");
    if (c->um->nbface > 0) MDB_LAUNCH(c, k_um_boundary, cdiv(c->um->nbface, 128), 128, 0, m, S, O, c->um->d_bface, c->um->nbface, weight, c->time, d_r);
    c->phase_end();
    c->phase_begin("residual_coupling");
    for (int q = 0; q < O.ncp; ++q)
      if (go.faces[q].n > 0)
        MDB_LAUNCH(c, k_um_coupling_residual, cdiv(go.faces[q].n, 128), 128, 0, m, S, O.cp[q], go.faces[q].d_cell, go.faces[q].d_face, go.faces[q].n, weight, d_x, d_r);
    c->phase_end();
  };
  // every update of r is a reduction: the boundary and coupling faces run beside the cells on the auxiliary stream
  const bool fork = !serial && !by_rows;
  if (fork) {
    MDB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    StreamScope aux(c, c->aux_stream);
    MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
    faces();
    MDB_CUDA(cudaEventRecord(c->ev_join, c->stream));
  }
  c->phase_begin("residual_volume");
  UMLists L;
  if (!by_rows && !by_cells && um_lists(c, go, nullptr, L)) {
    if (L.start[L.nlist] > 0) MDB_LAUNCH(c, k_um_residual_lists, cdiv(L.start[L.nlist], 256), 256, 0, m, L, weight, c->time, d_x, d_r);
  } else if (by_rows) {
    for (int k = 0; k < S.nchild; ++k)
      if (S.size[k] > 0) MDB_LAUNCH(c, k_um_residual_rows, cdiv(S.size[k], 128), 128, 0, m, S, O, k, weight, c->time, d_x, d_r);
  } else if (O.nsp > 0) MDB_LAUNCH(c, k_um_residual_cells, cdiv(m.ne, 128), 128, 0, m, S, O, weight, c->time, d_x, d_r);
  c->phase_end();
  if (fork) MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0)); else faces();
  c->phase_begin("residual_constrain");
  constrain_residual(c, d_r);
  c->phase_end();
}

void um_jacobian(Ctx* c, const GridOp& go, Matrix& M, double weight, const double* d_x, bool overwrite)
{
  const UMDev m = um_dev(c);
  const UMSpace S = um_space(c);
  const UMOps O = um_ops(c, go.sps, go.cps);
  c->phase_begin("jacobian_rows");
  static const bool simple = getenv("MDB_UM_JAC_SIMPLE") != nullptr;      // A/B switch: the unstaged row kernel
  static const bool staged = getenv("MDB_UM_JAC_STAGED") != nullptr;      // A/B switch: the owner-computes staged row kernel (fixed summation order)
  static const bool cells = getenv("MDB_UM_JAC_CELLS") != nullptr;        // A/B switch: cell-wise kernel with reductions into precomputed slots
  UMLists L;
  UMPairPlan* plan = (!simple && !staged && !cells && O.nsp > 0) ? um_pair_plan(c, go, M) : nullptr;
  const bool cellwise = !plan && !simple && !staged && um_lists(c, go, &M, L);
  if (plan) {
    const UMOps Ov = um_ops(c, go.sps, std::vector<int>());
    MDB_LAUNCH(c, k_um_jacobian_pairs, cdiv(c->ndof, UM_PROWS), UM_PROWS, 0, m, Ov, c->ndof, plan->rec, plan->wptr, weight, overwrite ? 1 : 0, M.d_rowptr, M.d_val);
  }
  static const bool serial = getenv("MDB_UM_JAC_SERIAL") != nullptr;      // A/B switch: coupling kernel after the volume kernel
  bool forked = false;
  if (cellwise) {
    if (overwrite) MDB_CUDA(cudaMemsetAsync(M.d_val, 0, M.nnz * sizeof(double), c->stream));
    if (!serial && O.ncp > 0) {
      // every update of the values is a reduction: the coupling faces run beside the cells on the auxiliary stream
      MDB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
      StreamScope aux(c, c->aux_stream);
      MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
      for (int q = 0; q < O.ncp; ++q)
        if (go.faces[q].n > 0)
          MDB_LAUNCH(c, k_um_coupling_jacobian, cdiv(6 * go.faces[q].n, 64), 64, 0, m, S, O.cp[q], go.faces[q].d_cell, go.faces[q].d_face, go.faces[q].n, weight, d_x,
                     M.d_rowptr, M.d_colidx, M.d_val);
      MDB_CUDA(cudaEventRecord(c->ev_join, c->stream));
      forked = true;
    }
    if (L.start[L.nlist] > 0) MDB_LAUNCH(c, k_um_jacobian_cells, cdiv(L.start[L.nlist], 256), 256, 0, m, L, weight, M.d_val);
  }
  for (int k = 0; k < S.nchild && !cellwise && !plan; ++k) {
    if (S.size[k] == 0) continue;
    auto it = M.slot_tables.find(std::make_pair(k, -1));
    if (simple || it == M.slot_tables.end())
      MDB_LAUNCH(c, k_um_jacobian_rows, cdiv(S.size[k], 128), 128, 0, m, S, O, k, weight, overwrite ? 1 : 0, M.d_rowptr, M.d_colidx, M.d_val);
    else
      MDB_LAUNCH(c, k_um_jacobian_staged, cdiv(S.size[k], UM_JROWS), UM_JROWS, 0, m, S, O, k, weight, overwrite ? 1 : 0, M.d_rowptr, it->second, M.d_val);
  }
  c->phase_end();
  if (forked) MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  c->phase_begin("jacobian_coupling");
  for (int q = 0; q < O.ncp && !forked; ++q)
    if (go.faces[q].n > 0)
      MDB_LAUNCH(c, k_um_coupling_jacobian, cdiv(6 * go.faces[q].n, 64), 64, 0, m, S, O.cp[q], go.faces[q].d_cell, go.faces[q].d_face, go.faces[q].n, weight, d_x,
                 M.d_rowptr, M.d_colidx, M.d_val);
  c->phase_end();
  c->phase_begin("jacobian_dirichlet");
  if (c->ncon > 0) MDB_LAUNCH(c, k_um_dirichlet_rows, cdiv(c->ncon, 128), 128, 0, c->d_conlist, c->ncon, M.d_rowptr, M.d_diag, M.d_val);
  c->phase_end();
}

} // namespace mdb
