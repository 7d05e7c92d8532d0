// common.cuh — shared host/device definitions of libmdb200 (sm_100a only).
//
// Data layout in HBM (all device arrays are owned by the context):
//   cell_sd      uint8  [ncells]            subdomain bit set of every cell (MultiDomainGrid marking)
//   lat2dof_k    int32  [nlat_k]            child k: lattice point -> index inside the child block, -1 if absent
//   dof2lat_k    int32  [size_k]            inverse map
//   constrained  uint8  [ndof]              Dirichlet flags; conlist int32 [ncon] the same as a compact list
//   CSR          int64 rowptr[ndof+1], int32 colidx[nnz], double val[nnz], uint32 rowinfo[ndof]
//   rowinfo      bits 0..26: which of the 3^d own-block stencil offsets exist in the row's pattern
//                bits 27..31: number of columns in the row that lie below the own child block (31 = search)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

#include "../../include/mdb200.h"

#define MDB_MAX_CHILDREN 8
#define MDB_MAX_SP_PER_CHILD 4
#define MDB_MAXLOC 27

namespace mdb {

// ------------------------------------------------------------------------------------------------
// error handling: exceptions inside the library, converted to codes at the C ABI
// ------------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MDB_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      throw mdb::Error(MDB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                      std::to_string(__LINE__) + ")");                                   \
  } while (0)

#define MDB_REQUIRE(cond, code, msg)                                                                     \
  do {                                                                                                   \
    if (!(cond)) throw mdb::Error(code, std::string(msg));                                               \
  } while (0)

// ------------------------------------------------------------------------------------------------
// structured mesh descriptor, passed by value to kernels
// ------------------------------------------------------------------------------------------------
struct MeshDesc {
  int dim;
  int n[3];              // local cells per axis
  double lower[3];       // local lower corner
  double h[3];           // (upper-lower)/n  of the local box
  // partition (multi-GPU): the local box is cells [off, off+n) of a global box of gn cells
  int gn[3];
  int off[3];
  double glower[3];      // == lower - off*h conceptually; kept separately so that analytic data see global coordinates
  int partitioned;

  __host__ __device__ int64_t ncells() const { return (int64_t)n[0] * n[1] * n[2]; }
  // coordinate of vertex plane i (local index) along axis d.  [UPSTREAM] YaspGrid equidistant coordinates: i*h.
  // With a partition the global plane index is used so that every rank computes bit-identical coordinates.
  __host__ __device__ double coord(int d, int i) const { return glower[d] + (double)(i + off[d]) * h[d]; }
  __host__ __device__ bool on_global_lower(int d, int i) const { return i + off[d] == 0; }
  __host__ __device__ bool on_global_upper(int d, int i) const { return i + off[d] == gn[d]; }
};

// device-side copies of the analytic data descriptors -------------------------------------------------
__host__ __device__ inline double eval_function(const mdb_function& f, int dim, const double* x, double t)
{
  switch (f.kind) {
  case MDB_FN_ZERO: return 0.0;
  case MDB_FN_CONST: return f.p[0];
  case MDB_FN_GAUSS: {
    double n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = f.p[1 + d] - x[d]; n2 += c * c; }
    return exp(-(f.p[0] * n2));
  }
  case MDB_FN_BOX: {
    for (int d = 0; d < dim; ++d) if (!(x[d] > f.p[1] && x[d] < f.p[2])) return 0.0;
    return f.p[0];
  }
  case MDB_FN_TESTPOISSON_J: {
    if (x[1] < 1E-6 || x[1] > 1.0 - 1E-6) return 0.0;
    if (x[0] > 1.0 - 1E-6 && x[1] > 0.5 + 1E-6) return -5.0;
    return 0.0;
  }
  case MDB_FN_INSTAT_G: {
    double n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = 0.5 - x[d]; n2 += c * c; }
    return f.p[0] * sin(f.p[1] * t) * exp(-n2);
  }
  case MDB_FN_POLY2: {
    double v = f.p[0], n2 = 0.0;
    for (int d = 0; d < dim; ++d) { v += f.p[1 + d] * x[d]; n2 += x[d] * x[d]; }
    return v + f.p[4] * n2;
  }
  case MDB_FN_DN_SOURCE: {
    bool box = true;
    for (int d = 0; d < dim; ++d) box = box && x[d] > 0.25 && x[d] < 0.375;
    if (box) return 50.0;
    double m2 = 0.0, n2 = 0.0;
    for (int d = 0; d < dim; ++d) { double c = 0.5 - x[d]; m2 += c * c; n2 += x[d] * x[d]; }
    if (sqrt(m2) < 0.2) return 60 * exp(-10 * n2);
    return 0.0;
  }
  case MDB_FN_DARCY_G: {                  // test/stokesdarcy2.cc:355-361
    const double scale = x[1] / f.p[2];
    return scale * f.p[0] + (1.0 - scale) * f.p[1];
  }
  case MDB_FN_DARCY_J:                    // :364-371
    return x[1] < 1E-6 ? f.p[0] * x[0] * (f.p[1] - x[0]) : 0.0;
  case MDB_FN_PRESSURE_DROP: {            // :191-203
    double y;
    if (x[0] < 1E-6) y = x[1] > f.p[3] ? f.p[0] : -f.p[1];
    else if (x[0] > f.p[2] - 1E-6) y = x[1] > f.p[3] ? -f.p[1] : f.p[0];
    else y = 0.0;
    return y * (f.p[4] - x[1]);
  }
  }
  return 0.0;
}

// 0 = dirichlet, 1 = neumann, 2 = none
__host__ __device__ inline int eval_bctype(const mdb_bctype& b, int dim, const double* xg, bool is_boundary)
{
  (void)dim;
  switch (b.kind) {
  case MDB_BC_ALL_DIRICHLET: return is_boundary ? 0 : 2;
  case MDB_BC_ALL_NEUMANN: return is_boundary ? 1 : 2;
  case MDB_BC_TESTPOISSON:
    if (!is_boundary) return 2;
    if (xg[1] < 1E-6 || xg[1] > 1.0 - 1E-6) return 1;
    if (xg[0] > 1.0 - 1E-6 && xg[1] > 0.5 + 1E-6) return 1;
    return 0;
  case MDB_BC_DARCY_RIGHT:                // test/stokesdarcy2.cc:338-353; Outflow (b = 0) acts like Neumann with o = j
    if (!is_boundary) return 2;
    if (xg[1] < 1E-6 && xg[0] < b.p[0]) return 1;
    if (xg[0] > b.p[0] - 1E-6 && xg[1] < b.p[1]) return 0;
    return 1;
  }
  return 2;
}

__host__ __device__ inline bool cond_applies(int cond_kind, uint32_t mask, uint32_t sd)
{
  return cond_kind == MDB_COND_EQUALITY ? (sd == mask) : ((sd & mask) == mask);
}

// Gauss-Legendre rules on [0,1]; order p -> ceil((p+1)/2) points.  Evaluated on the host with correctly rounded
// sqrt/div so that every consumer sees the same bits.
struct Rule1D { int m; double x[5]; double w[5]; };
Rule1D gauss1d(int order);

// ------------------------------------------------------------------------------------------------
// host-side model
// ------------------------------------------------------------------------------------------------
// `subdomain` of a coupling (mortar) space: no cell carries its DOFs — it is bound per intersection (globalassembler.hh:212-231);
// bit 30 of a cell's subdomain byte is never set, so every "child lives on this cell" test of the volume kernels fails
#define MDB_SUBDOMAIN_IFACE 30

struct Child {
  int fe_kind = 0, k = 1, subdomain = -1;
  // QkDG (MDB_FE_DGQk): every coefficient belongs to its cell.  The child keeps the lattice bookkeeping of the conforming spaces
  // with lattice stride k + 1 instead of k (cell c, local index a <-> lattice point (k+1) c + a: no point is shared), numbered
  // cell block after cell block: DOF = (index of the cell in the subdomain) * (k+1)^dim + local index
  bool dg = false;
  int stride() const { return dg ? k + 1 : k; }
  // CouplingGridFunctionSpace (couplinggridfunctionspace.hh:45-249): lives on the intersections whose inside cell satisfies
  // the left condition and whose outside cell satisfies the right one (SubProblemSubProblemInterface :14-44)
  bool iface = false; int lkind = 0, rkind = 0; uint32_t lmask = 0, rmask = 0;
  int lat_n[3] = {1, 1, 1};
  int64_t nlat = 0;
  int32_t* d_lat2dof = nullptr;
  int32_t* d_dof2lat = nullptr;
  int64_t size = 0, offset = 0;
  int lat_lo[3] = {0, 0, 0}, lat_hi[3] = {0, 0, 0};   // bounding box (inclusive) of the lattice points that carry a DOF
  int nloc(int dim) const { int r = 1; for (int d = 0; d < dim; ++d) r *= (k + 1); return r; }
  // generic spaces on a triangle mesh (ufe.cu): P2 keeps vertex DOFs (d_lat2dof, nvs of them) followed by edge DOFs (d_edge2dof);
  // an OPB space numbers the subdomain's cells (d_cell2dof), um_nloc coefficients per cell
  int32_t* d_edge2dof = nullptr; int32_t* d_cell2dof = nullptr; int64_t nvs = 0; int um_nloc = 3;
};

struct SubProblemDesc {
  int op_id = 0;
  mdb_poisson_params poisson{};
  mdb_l2_params l2{};
  mdb_convdiffdg_params cddg{};
  mdb_taylorhood_params th{};
  mdb_darcy_params darcy{};
  int32_t* d_um_vtx = nullptr; int32_t* d_um_dof = nullptr; int64_t um_ncells = 0;   // unstructured.cu: [3][um_ncells] vertex ids / container indices of the cells the subproblem applies to (Morton order)
  int32_t* d_lmap = nullptr; int nloc = 0;     // ufe.cu: [ncells][nloc] container indices of the subproblem's local space (-1: not on the cell)
  int dg_k = 1;                // MDB_OP_CONVDIFF_DG: degree of the child space (quadrature order = intorderadd + 2 k)
  int cond_kind = 0; uint32_t mask = 0;
  std::vector<int> children;
  bool doPatternVolume() const { return true; }
  bool doPatternSkeleton() const { return op_id == MDB_OP_CONVDIFF_DG; }
  bool doAlphaVolume() const { return true; }
  bool doLambdaVolume() const { return op_id == MDB_OP_POISSON; }
  bool doLambdaBoundary() const { return op_id == MDB_OP_POISSON; }
  int quad_order() const { return op_id == MDB_OP_POISSON ? poisson.quad_order : op_id == MDB_OP_CONVDIFF_DG ? cddg.intorderadd + 2 * dg_k : l2.quad_order; }
};

struct CouplingDesc {
  int op_id = 0;
  mdb_cvcf_params cvcf{};
  mdb_propflow_params propflow{};
  mdb_mortar_params mortar{};
  mdb_ndcoupling_params nd{};
  mdb_stokesdarcy_params sd{};
  int local_sp = -1, remote_sp = -1, jac_mode = 0;
  int coupling_child = -1;       // >= 0: EnrichedCoupling (coupling.hh:123-192) with this coupling space
  int ncomp = 1;                 // component blocks of a power coupling space: children coupling_child .. coupling_child + ncomp - 1
  bool enriched() const { return coupling_child >= 0; }
};

struct FaceList {      // interface faces on which one coupling fires, in traversal order
  int32_t* d_cell = nullptr;     // inside cell (local index)
  int8_t* d_face = nullptr;      // face number 0..2*dim-1
  int64_t n = 0;
};

// Unstructured simplex mesh (2-D triangles; unstructured.cu): the mesh a grid manager such as UGGrid hands to the reference's
// drivers after GmshReader (test/stokesdarcy2.cc:575-604).  Cell and vertex indices are the caller's ([UPSTREAM] leaf index order).
struct UMesh {
  int64_t nv = 0, ne = 0;
  double* d_xy = nullptr;        // [nv][2] vertex coordinates
  int32_t* d_conn = nullptr;     // [ne][3] vertex indices, local vertex i <-> P1 shape function i
  int32_t* d_nbr = nullptr;      // [ne][3] neighbour over face f (DUNE simplex numbering: face 0 = {0,1}, 1 = {0,2}, 2 = {1,2}); -1 = boundary
  int32_t* d_v2c_ptr = nullptr;  // [nv+1] vertex -> incident cells
  int32_t* d_v2c = nullptr;      // [3 ne] 4 * cell + local vertex index, cells ascending per vertex
  int32_t* d_bface = nullptr;    // [nbface] 3 * cell + face of the faces without a neighbour (the physical boundary), ascending
  int64_t nbface = 0;
  int64_t nedges = 0;
  int32_t* d_cell_edge = nullptr; // [ne][3] edge index of face f: order of first appearance in (cell, local face) order
  int32_t* d_group = nullptr;    // [ne] physical group (mdb_cell_groups), zero by default
  int32_t* d_cell_order = nullptr; // [ne] cells in Morton order of their centroids: the processing order of the cell-wise kernels
  void* ufe_tables = nullptr;    // ufe.cu: device tables of the generic finite elements
};

struct GridOp {
  std::vector<int> sps, cps;
  std::vector<FaceList> faces;   // one per coupling
  // sorted list of the DOFs whose coefficients the Jacobian kernels of this operator read (the coupling faces' cell DOFs;
  // the volume operators of the closed set are linear).  Built on first use by mdb_jacobian with a host vector.
  int32_t* d_xneed = nullptr; int64_t n_xneed = -1;
  // residual_volume plan per child (assemble.cu): 1 = row taken by the bulk-copy stencil kernel; the other rows as a list
  uint8_t* d_res_fast[MDB_MAX_CHILDREN] = {nullptr}; int32_t* d_res_rest[MDB_MAX_CHILDREN] = {nullptr};
  int64_t n_res_rest[MDB_MAX_CHILDREN] = {-1, -1, -1, -1, -1, -1, -1, -1};
};

struct FdPlan;                      // fdcoupling.cu: tables of the register-variant FD coupling kernel on a planar interface
void fd_plan_free(FdPlan* p);

struct Matrix {
  std::vector<int> gos;
  int64_t nnz = 0;
  int max_row = 0;
  int64_t* d_rowptr = nullptr;
  int32_t* d_colidx = nullptr;
  double* d_val = nullptr;
  uint32_t* d_rowinfo = nullptr;
  int32_t* d_diag = nullptr;     // offset of the diagonal inside the row
  // per child block: value-array span of its rows, and the compact list of its irregular rows (stencil mask not full)
  int64_t child_begin[MDB_MAX_CHILDREN] = {0}, child_end[MDB_MAX_CHILDREN] = {0};
  int32_t* d_irregular[MDB_MAX_CHILDREN] = {nullptr};
  int64_t n_irregular[MDB_MAX_CHILDREN] = {0};
  // static records of the irregular rows (assemble.cu: k_rowrec_build / k_jacobian_rows_rec), built on first use:
  // value-array start, stencil mask | columns below the block, length | validity of the 2^d incident cells, their subdomain bytes
  int64_t* d_rr_rs[MDB_MAX_CHILDREN] = {nullptr}; uint32_t* d_rr_mask[MDB_MAX_CHILDREN] = {nullptr}; uint32_t* d_rr_meta[MDB_MAX_CHILDREN] = {nullptr};
  unsigned long long* d_rr_sd[MDB_MAX_CHILDREN] = {nullptr};
  int rr_state[MDB_MAX_CHILDREN] = {0};          // 0 not built, 1 usable, -1 a row cannot be staged: gather kernel
  int32_t* d_seg_start[MDB_MAX_CHILDREN] = {nullptr};   // runs of rows with equal simple/irregular status (n_seg + 1 entries)
  int64_t n_seg[MDB_MAX_CHILDREN] = {0};
  longlong2* d_seg_span[MDB_MAX_CHILDREN] = {nullptr};  // value-array span [x, y) of every run (empty for irregular runs)
  // per (grid operator, coupling): uint8 offsets inside the row of every entry a coupling face touches
  std::map<std::pair<int, int>, uint8_t*> slot_tables;
  std::map<std::pair<int, int>, FdPlan*> fd_plans;
  // preconditioner caches
  double* d_dinv = nullptr;
  bool dinv_valid = false; int dinv_precond = -1;
  // sequential preconditioners (precond.cu): ready flags / row counter of the sync-free sweeps, ILU(0) factors
  int* d_ready = nullptr; unsigned long long* d_sweep_counter = nullptr; int sweep_epoch = 0; int32_t* d_sweep_perm = nullptr; ulonglong2* d_vrec = nullptr;
  double* d_ilu = nullptr; bool ilu_valid = false;
  int32_t* d_cell_perm = nullptr; int64_t sweep_cells = 0; int sweep_cell_img = 0;   // all-DG matrices: cells in wavefront order (first row of each), image doubles per warp
  void* direct = nullptr;            // direct.cu: band factorisation cache
  void* mg = nullptr;                // mg.cu: multigrid hierarchy (MDB_PRECOND_MG)
  int dir_early_version = -1; bool dir_early_ok = false;   // assemble.cu: no constrained row is a streamed (simple) row for constraint version dir_early_version
  std::vector<std::pair<int, double>> asm_log;   // (grid operator, weight) of the mdb_jacobian calls since the values were last zeroed
  void* qk_plans[MDB_MAX_CHILDREN] = {nullptr};   // generic.cu: classes / runs of the simple rows of a Qk (k > 1) child, built on first use
  std::map<std::vector<int>, void*> um_pair_plans;   // unstructured.cu: pair records of the owner-computes Jacobian kernel per subproblem list
  // SpMV: ascending list of the rows the stencil kernel leaves to the row-list kernel (solver.cu); -1 = not built yet
  int32_t* d_spmv_rest = nullptr; int64_t n_spmv_rest = -1;
  // shift-run product (solver.cu, k_spmv_runs): chunks of <= 32 consecutive rows whose column lists are those of the row before
  // shifted by one; -1 = not analysed yet, 0 = the matrix has too few such rows (vector-CSR kernel)
  int4* d_sr_chunks = nullptr; int64_t n_sr_chunks = -1; int sr_maxm = 0;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // deferred per-phase device timing: event pairs recorded on the stream without host synchronisation and resolved
  // when mdb_phase_ms() is called
  struct PhaseRec { std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev; size_t used = 0; double last_ms = -1.0; };
  std::map<std::string, PhaseRec> phases;
  PhaseRec* cur_phase = nullptr; bool nvtx_open = false;
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // second stream for host->device traffic that overlaps assembly kernels; byte counters of the host<->device copies
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy_done = nullptr, ev_stage_free = nullptr;
  // auxiliary compute stream: mdb_jacobian runs the irregular rows + FD coupling there while the simple rows stream out
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int64_t h2d_bytes = 0, d2h_bytes = 0;

  MeshDesc mesh{};
  UMesh* um = nullptr;               // non-null: unstructured mesh (mesh.dim = 2, mesh.n = {ne, 1, 1} so that ncells() holds)
  bool ufe = false;                  // unstructured mesh with spaces / operators beyond P1 Poisson: the generic element path (ufe.cu)
  bool have_mesh = false, have_sd = false, finalized = false;
  uint8_t* d_cell_sd = nullptr;
  std::vector<Child> children;
  std::vector<SubProblemDesc> sps;
  std::vector<CouplingDesc> cps;
  std::vector<GridOp> gos;
  std::vector<Matrix> mats;
  int64_t ndof = 0;
  uint8_t* d_constrained = nullptr;
  int32_t* d_conlist = nullptr;
  int64_t ncon = 0;
  int con_version = 0;               // bumped whenever the constraint set changes (rebuild_conlist)
  double time = 0.0;
  std::vector<int> con_sps; std::vector<mdb_bctype> con_bcs; bool con_recorded = true;   // the last mdb_constraints_assemble call (mg.cu replays it on the coarse levels)
  int ssor_sweeps = 3;               // SeqSSOR(A, n, 1.0): ISTLBackend_SEQ_BCGS_SSOR uses n = 3 (test/testpoisson.cc:291)

  // scratch
  void* d_scratch = nullptr; size_t scratch_bytes = 0;
  double* d_Ae = nullptr;            // element matrix table
  std::map<int, double*> mortar_tabs; // mortar.cu: shape-function tables per enriched coupling (index into cps)
  double* d_ftab = nullptr;          // face shape-function tables of the coupling kernels
  double* d_ftab_g = nullptr;        // the same for Qk x Qk couplings (generic.cu)
  double* d_dgtab = nullptr;         // skeleton / boundary block tables of the QkDG kernels (dg.cu)
  // staging for host<->device vectors
  double* d_stage[4] = {nullptr, nullptr, nullptr, nullptr};
  // solver workspace
  double* d_work = nullptr; int64_t work_n = 0; int work_vecs = 0;
  double* d_onestep = nullptr; int64_t onestep_n = 0;      // u, R, z of mdb_onestep_apply
  double* d_scal = nullptr;          // device scalars
  unsigned int* d_fill_ctr = nullptr; // run counters of the hybrid fill kernel (one per child job)
  double* d_partial = nullptr;       // block partial sums
  double* h_pinned = nullptr;        // pinned host scalars

  // multi-GPU (multigpu.cu)
  void* sring = nullptr; bool sring_borrowed = false;          // scalar windows of all ranks (multigpu.cu: peer-memory all-reduce of the Krylov scalars); coarse multigrid levels share the fine context's
  void* nccl_comm = nullptr; int nranks = 1, rank = 0; bool comm_borrowed = false;   // borrowed: a coarse multigrid level shares the fine context's communicator
  int part_axis = -1;                // cut axis of the slab partition (-1: not partitioned)
  uint8_t* d_owned = nullptr;        // 1 = row owned by this rank (null: all owned)
  int64_t owned_rows = 0;
  std::vector<int64_t> child_owned;  // owned rows per child block
  double* d_halo_recv = nullptr; int64_t halo_recv_n = 0;       // receive window [parity][side][halo_cap] + sequence numbers
  int64_t halo_cap = 0;              // doubles per (parity, side) slot
  unsigned long long halo_seq = 0;   // exchanges issued so far (same on every rank: SPMD)
  unsigned int* d_halo_counter = nullptr;                       // [0] block counter of the push kernel, [2] timeout flag
  double* peer_lower = nullptr; double* peer_upper = nullptr;   // mapped windows of the neighbours
  std::vector<void*> ipc_opened;
  int lower_rank = -1, upper_rank = -1;
  void* halo_lists = nullptr;        // multigpu.cu: index-list halo of an unstructured partition (mdb_partition_lists)
  double halo_timeout_s = 0.0;       // bound of the halo wait (mdb_set_halo_timeout; 0: MDB_HALO_TIMEOUT_S or 30 s)

  void* scratch(size_t bytes);
  void phase_begin(const char* name);
  void phase_end();
  double phase_mean_ms(const char* name, int* count);
  void phase_reset();
};

#define SR_MAXR 16
#define SR_MAXV 4
struct SrSpec { int nranks, rank; unsigned long long seq; double* win[SR_MAXR]; int* err; long long timeout_clocks; double* done; };
#ifdef __CUDACC__
// Window of a rank: double val[2][SR_MAXR][SR_MAXV] followed by unsigned long long flag[2][SR_MAXR].  Exchange number seq (the same on
// every rank) uses parity seq & 1: thread t < nranks stores this rank's values into rank t's window and publishes seq with a
// release store at system scope, then waits for rank t's flag in its own window; thread 0 adds the contributions in rank order, so
// every rank obtains the same bits.  Two parities suffice: a rank publishes seq + 1 only after it has read seq, and nobody starts
// seq + 2 before every flag of seq + 1 has arrived.  Call with all threads of ONE block; loc[] in shared memory holds the local
// values on entry and the sums on exit.
__device__ __forceinline__ void sr_exchange(const SrSpec& R, double* loc, int n)
{
  const int t = threadIdx.x, par = (int)(R.seq & 1);
  __shared__ int sr_ok;
  if (t == 0) sr_ok = 1;
  __syncthreads();
  if (t < R.nranks) {
    double* w = R.win[t];
    for (int k = 0; k < n; ++k) w[(par * SR_MAXR + R.rank) * SR_MAXV + k] = loc[k];
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(w + 2 * SR_MAXR * SR_MAXV) + par * SR_MAXR + R.rank), "l"(R.seq) : "memory");
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(R.win[R.rank] + 2 * SR_MAXR * SR_MAXV) + par * SR_MAXR + t;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= R.seq) break;
      if (clock64() - t0 > R.timeout_clocks) { sr_ok = 0; break; }
    }
  }
  __syncthreads();
  if (t == 0) {
    if (!sr_ok) { atomicExch(R.err, 1); if (R.done) *R.done = 1.0; }
    const double* mine = R.win[R.rank];
    for (int k = 0; k < n; ++k) {
      double a = 0.0;
      for (int r = 0; r < R.nranks; ++r) a += __ldcv(mine + (par * SR_MAXR + r) * SR_MAXV + k);
      loc[k] = a;
    }
  }
  __syncthreads();
}
#endif

// run a section of host code with another stream as the context's launch stream
struct StreamScope {
  Ctx* c; cudaStream_t save;
  StreamScope(Ctx* c_, cudaStream_t s) : c(c_), save(c_->stream) { c->stream = s; }
  ~StreamScope() { c->stream = save; }
};

inline bool is_device_ptr(const void* p)
{
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// RAII view of a caller vector on the device: device pointers pass through, host pointers are staged.
struct VecIn {
  const double* d = nullptr;
  VecIn(Ctx* c, const double* p, int64_t n, int slot);
};
struct VecInOut {
  Ctx* c; double* d = nullptr; double* host = nullptr; int64_t n;
  VecInOut(Ctx* c, double* p, int64_t n, int slot, bool load);
  void commit();     // copy back to the host buffer if staged
};

// module entry points (implemented in the .cu files)
void build_dofmap(Ctx* c);
void assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n);
void rebuild_conlist(Ctx* c);
void interpolate(Ctx* c, const int* sps, const mdb_function* fns, int n, double t, double* d_x);
void export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs);
void build_facelists(Ctx* c, GridOp& go);
void build_xneed(Ctx* c, GridOp& go);
void build_pattern(Ctx* c, Matrix& m);
void residual_volume(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r);
bool residual_can_split(Ctx* c, const GridOp& go);
void residual_prepare(Ctx* c, const GridOp& go);
void residual_volume_part(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r, int part);
void residual_boundary(Ctx* c, const GridOp& go, double weight, double* d_r);
void residual_coupling(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r);
void constrain_residual(Ctx* c, double* d_r);
void launch_element_matrices(Ctx* c, const std::vector<int>& sps, double weight);
void jacobian_volume(Ctx* c, const GridOp& go, Matrix& m, double weight, bool overwrite, int part);
void jacobian_coupling(Ctx* c, const GridOp& go, Matrix& m, double weight, const double* d_x, bool late);
// mortar.cu: EnrichedCoupling with MortarPoissonCoupling (test/testmortarpoisson.cc:117-250)
void residual_enriched(Ctx* c, const CouplingDesc& cp, const FaceList& fl, double weight, const double* d_x, double* d_r);
void jacobian_enriched(Ctx* c, const CouplingDesc& cp, const FaceList& fl, Matrix& m, double weight, const double* d_x, std::pair<int, int> key);
// nd.cu: ConvectionDiffusionDGNeumannDirichletCoupling (test/neumanndirichletcoupling.hh)
void residual_nd(Ctx* c, const CouplingDesc& cp, const FaceList& fl, double weight, const double* d_x, double* d_r);
void jacobian_nd(Ctx* c, const CouplingDesc& cp, const FaceList& fl, Matrix& m, double weight);
void mark_iface_present(Ctx* c, const Child& ch, uint8_t* d_present);
// dg.cu: QkDG children with ConvectionDiffusionDG (volume + skeleton + boundary terms, owner-computes)
void jacobian_dg(Ctx* c, const GridOp& go, Matrix& m, int child, double weight, bool overwrite);
void residual_dg(Ctx* c, const GridOp& go, int child, double weight, const double* d_x, double* d_r);
void jacobian_volume_generic(Ctx* c, const GridOp& go, Matrix& m, int child, bool overwrite);      // generic.cu: Qk children, k > 1
void residual_volume_generic(Ctx* c, const GridOp& go, int child, const double* d_x, double* d_r);
void dirichlet_rows(Ctx* c, Matrix& m);
bool dirichlet_rows_off_the_fill(Ctx* c, Matrix& m);      // true: no constrained row is written by the streamed fill (the pass may run beside it)
void spmv(Ctx* c, const Matrix& m, const double* d_x, double* d_y);
void spmv_partitioned(Ctx* c, const Matrix& m, double* d_x, double* d_y);   // halo exchange overlapped with the interior rows
int solve(Ctx* c, Matrix& m, int method, int precond, double reduction, int maxit, int fixed_iters,
          const double* d_rhs, double* d_z, mdb_solve_stats* stats);
void setup_partition(Ctx* c);                       // multigpu.cu: ownership flags + halo window of a partitioned mesh
void allreduce_scalars(Ctx* c, double* d_vals, int n);
// Peer-memory all-reduce of up to SR_MAXV scalars (multigpu.cu).  SrSpec is passed by value to the kernel that ends in the exchange.
bool sring_connected(const Ctx* c);
SrSpec sring_next(Ctx* c, double* d_done);          // the descriptor of the next exchange (advances the sequence number)
void sring_release(Ctx* c);
void halo_exchange(Ctx* c, double* d_x, double* d_done = nullptr);   // d_done: the solver's S_DONE slot, set when the wait times out
void halo_push(Ctx* c, const double* d_x);
void halo_wait(Ctx* c, double* d_x, double* d_done = nullptr);
void nccl_comm_release(Ctx* c);
void halo_lists_free(Ctx* c);
bool halo_timed_out(Ctx* c);
// unstructured.cu: the same steps on an unstructured triangle mesh with P1 spaces
void um_ingest(Ctx* c, int64_t nv, const double* xy, int64_t ne, const int64_t* conn);
void um_free(Ctx* c);
void um_pair_plan_free(void* plan);
void qk_plan_free(void* plan);
void um_build_dofmap(Ctx* c);
void um_export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs);
void um_assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n);
void um_interpolate(Ctx* c, const int* sps, const mdb_function* fns, int n, double t, double* d_x);
void um_build_facelists(Ctx* c, GridOp& go);
void um_build_pattern(Ctx* c, Matrix& m);
void um_residual(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r);
void um_jacobian(Ctx* c, const GridOp& go, Matrix& m, double weight, const double* d_x, bool overwrite);
void precond_apply(Ctx* c, Matrix& m, int kind, int sweeps, const double* d_d, double* d_v);   // precond.cu: SSOR / ILU0
// ufe.cu: generic finite elements on a triangle mesh (P1 / P2 / orthonormal Pk, multi-child subproblems): the Stokes-Darcy path
void ufe_free(Ctx* c);
void ufe_build_dofmap(Ctx* c);
void ufe_export_dofmap(Ctx* c, int64_t* cell_ptr, int64_t* cell_dofs);
void ufe_assemble_constraints(Ctx* c, const int* sps, const mdb_bctype* bcs, int n);
void ufe_build_facelists(Ctx* c, GridOp& go);
void ufe_build_pattern(Ctx* c, Matrix& m);
void ufe_residual(Ctx* c, const GridOp& go, double weight, const double* d_x, double* d_r);
void ufe_jacobian(Ctx* c, const GridOp& go, Matrix& m, double weight, const double* d_x, bool overwrite);
// direct.cu: banded LU with partial pivoting (MDB_SOLVER_DIRECT)
int direct_solve(Ctx* c, Matrix& m, const double* d_b, double* d_z, mdb_solve_stats* stats);
void direct_free(Matrix& m);
void direct_invalidate(Matrix& m);
// mg.cu: geometric multigrid V-cycle (MDB_PRECOND_MG)
void mg_apply(Ctx* c, Matrix& m, const double* d_d, double* d_v);
void mg_free(Matrix& m);
int mg_levels(const Matrix& m);
int mg_setup(Ctx* c, Matrix& m);
::mdb_ctx* mg_level_ctx(Matrix& m, int level);

template <typename T> inline T* dalloc(size_t n)
{
  T* p = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&p, n * sizeof(T));
  if (e != cudaSuccess) throw Error(MDB_ENOMEM, std::string("cudaMalloc of ") + std::to_string(n * sizeof(T)) + " bytes: " + cudaGetErrorString(e));
  return p;
}
template <typename T> inline void dfree(T*& p) { if (p) cudaFree(p); p = nullptr; }

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#define MDB_LAUNCH(c, kernel, grid, block, smem, ...)                           \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (c)->stream>>>(__VA_ARGS__);              \
    (c)->launches++;                                                            \
    MDB_CUDA(cudaGetLastError());                                               \
  } while (0)

} // namespace mdb
