// capi.cu — the C ABI of libmdb200.so (include/mdb200.h): context, mesh, spaces, participants, entry points.
#include "common.cuh"
#include <nvtx3/nvToolsExt.h>
#include <cmath>

namespace mdb {

static std::string g_create_error;

Rule1D gauss1d(int order)
{
  Rule1D r;
  int m = (order + 2) / 2;
  if (m < 1) m = 1;
  MDB_REQUIRE(m <= 5, MDB_EINVAL, "quadrature order too high (max 9)");
  r.m = m;
  switch (m) {
  case 1: r.x[0] = 0.5; r.w[0] = 1.0; break;
  case 2: {
    double a = 0.5 / std::sqrt(3.0);
    r.x[0] = 0.5 - a; r.x[1] = 0.5 + a; r.w[0] = 0.5; r.w[1] = 0.5; break; }
  case 3: {
    double a = 0.5 * std::sqrt(3.0 / 5.0);
    r.x[0] = 0.5 - a; r.x[1] = 0.5; r.x[2] = 0.5 + a;
    r.w[0] = 5.0 / 18.0; r.w[1] = 8.0 / 18.0; r.w[2] = 5.0 / 18.0; break; }
  case 4: {
    double s = std::sqrt(6.0 / 5.0);
    double a = 0.5 * std::sqrt(3.0 / 7.0 - 2.0 / 7.0 * s);
    double b = 0.5 * std::sqrt(3.0 / 7.0 + 2.0 / 7.0 * s);
    double wa = (18.0 + std::sqrt(30.0)) / 72.0;
    double wb = (18.0 - std::sqrt(30.0)) / 72.0;
    r.x[0] = 0.5 - b; r.x[1] = 0.5 - a; r.x[2] = 0.5 + a; r.x[3] = 0.5 + b;
    r.w[0] = wb; r.w[1] = wa; r.w[2] = wa; r.w[3] = wb; break; }
  case 5: {
    double s = 2.0 * std::sqrt(10.0 / 7.0);
    double a = 0.5 * (1.0 / 3.0) * std::sqrt(5.0 - s);
    double b = 0.5 * (1.0 / 3.0) * std::sqrt(5.0 + s);
    double wa = (322.0 + 13.0 * std::sqrt(70.0)) / 1800.0;
    double wb = (322.0 - 13.0 * std::sqrt(70.0)) / 1800.0;
    r.x[0] = 0.5 - b; r.x[1] = 0.5 - a; r.x[2] = 0.5; r.x[3] = 0.5 + a; r.x[4] = 0.5 + b;
    r.w[0] = wb; r.w[1] = wa; r.w[2] = 128.0 / 450.0; r.w[3] = wa; r.w[4] = wb; break; }
  }
  return r;
}

void* Ctx::scratch(size_t bytes)
{
  if (bytes > scratch_bytes) {
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr; scratch_bytes = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&d_scratch, want);
    if (e != cudaSuccess) throw Error(MDB_ENOMEM, "scratch cudaMalloc failed");
    scratch_bytes = want;
  }
  return d_scratch;
}

void Ctx::phase_begin(const char* name)
{
  PhaseRec& r = phases[name];
  if (nvtx_open) nvtxRangePop();
  nvtxRangePushA(name); nvtx_open = true;
  cur_phase = nullptr;
  if (r.used == r.ev.size()) {
    if (r.ev.size() >= 4096) return;            // pool full: stop recording this phase until the next reset
    cudaEvent_t a, b;
    MDB_CUDA(cudaEventCreate(&a)); MDB_CUDA(cudaEventCreate(&b));
    r.ev.emplace_back(a, b);
  }
  MDB_CUDA(cudaEventRecord(r.ev[r.used].first, stream));
  cur_phase = &r;
}
void Ctx::phase_end()
{
  if (nvtx_open) { nvtxRangePop(); nvtx_open = false; }
  if (!cur_phase) return;
  MDB_CUDA(cudaEventRecord(cur_phase->ev[cur_phase->used].second, stream));
  cur_phase->used++;
  cur_phase = nullptr;
}
double Ctx::phase_mean_ms(const char* name, int* count)
{
  auto it = phases.find(name);
  if (count) *count = 0;
  if (it == phases.end()) return -1.0;
  PhaseRec& r = it->second;
  if (r.used == 0) return r.last_ms;
  MDB_CUDA(cudaStreamSynchronize(stream));
  double sum = 0.0;
  for (size_t i = 0; i < r.used; ++i) { float ms = 0.f; MDB_CUDA(cudaEventElapsedTime(&ms, r.ev[i].first, r.ev[i].second)); sum += ms; }
  if (count) *count = (int)r.used;
  r.last_ms = sum / r.used;
  return r.last_ms;
}
void Ctx::phase_reset()
{
  for (auto& kv : phases) { if (kv.second.used) phase_mean_ms(kv.first.c_str(), nullptr); kv.second.used = 0; }
}

static bool is_pinned_host(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

VecIn::VecIn(Ctx* c, const double* p, int64_t n, int slot)
{
  if (is_device_ptr(p)) { d = p; return; }
  if (!c->d_stage[slot]) c->d_stage[slot] = dalloc<double>(c->ndof > n ? c->ndof : n);
  MDB_CUDA(cudaMemcpyAsync(c->d_stage[slot], p, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes += n * (int64_t)sizeof(double);
  d = c->d_stage[slot];
  // Host vectors are plain inputs: the caller may reuse the buffer as soon as the call returns.  A copy from pageable memory
  // is staged before cudaMemcpyAsync returns; a copy from page-locked memory is a real DMA, so wait for it here.
  if (is_pinned_host(p)) MDB_CUDA(cudaStreamSynchronize(c->stream));
}
VecInOut::VecInOut(Ctx* c_, double* p, int64_t n_, int slot, bool load) : c(c_), n(n_)
{
  if (is_device_ptr(p)) { d = p; return; }
  host = p;
  if (!c->d_stage[slot]) c->d_stage[slot] = dalloc<double>(c->ndof > n ? c->ndof : n);
  d = c->d_stage[slot];
  if (load) {
    MDB_CUDA(cudaMemcpyAsync(d, p, n * sizeof(double), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n * (int64_t)sizeof(double);
    if (is_pinned_host(p)) MDB_CUDA(cudaStreamSynchronize(c->stream));      // see VecIn
  }
}
void VecInOut::commit()
{
  if (host) {
    MDB_CUDA(cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h_bytes += n * (int64_t)sizeof(double);
  }
}

static void free_matrix(Matrix& m)
{
  dfree(m.d_rowptr); dfree(m.d_colidx); dfree(m.d_val); dfree(m.d_rowinfo); dfree(m.d_diag); dfree(m.d_dinv); dfree(m.d_spmv_rest); dfree(m.d_sr_chunks); dfree(m.d_ready); dfree(m.d_sweep_counter); dfree(m.d_sweep_perm); dfree(m.d_cell_perm); dfree(m.d_vrec); dfree(m.d_ilu);
  for (int i = 0; i < MDB_MAX_CHILDREN; ++i) { dfree(m.d_irregular[i]); dfree(m.d_seg_start[i]); dfree(m.d_seg_span[i]); dfree(m.d_rr_rs[i]); dfree(m.d_rr_mask[i]); dfree(m.d_rr_meta[i]); dfree(m.d_rr_sd[i]); }
  for (auto& kv : m.slot_tables) cudaFree(kv.second);
  m.slot_tables.clear();
  for (auto& kv : m.fd_plans) fd_plan_free(kv.second);
  m.fd_plans.clear();
  direct_free(m);
  mg_free(m);
  for (int i = 0; i < MDB_MAX_CHILDREN; ++i) { qk_plan_free(m.qk_plans[i]); m.qk_plans[i] = nullptr; }
  for (auto& kv : m.um_pair_plans) um_pair_plan_free(kv.second);
  m.um_pair_plans.clear();
}

static void reset_spaces(Ctx* c)
{
  for (auto& ch : c->children) { dfree(ch.d_lat2dof); dfree(ch.d_dof2lat); dfree(ch.d_edge2dof); dfree(ch.d_cell2dof); }
  for (auto& sp : c->sps) { dfree(sp.d_lmap); dfree(sp.d_um_vtx); dfree(sp.d_um_dof); sp.um_ncells = 0; }
  for (auto& go : c->gos) { for (auto& f : go.faces) { dfree(f.d_cell); dfree(f.d_face); } dfree(go.d_xneed); for (int i = 0; i < MDB_MAX_CHILDREN; ++i) { dfree(go.d_res_fast[i]); dfree(go.d_res_rest[i]); go.n_res_rest[i] = -1; } }
  for (auto& m : c->mats) free_matrix(m);
  c->gos.clear(); c->mats.clear();
  for (auto& kv : c->mortar_tabs) cudaFree(kv.second);
  c->mortar_tabs.clear();
  dfree(c->d_constrained); dfree(c->d_conlist);
  for (int i = 0; i < 4; ++i) dfree(c->d_stage[i]);
  dfree(c->d_work); c->work_n = 0; c->work_vecs = 0;
  dfree(c->d_onestep); c->onestep_n = 0;
  dfree(c->d_owned); c->owned_rows = 0; c->child_owned.clear(); c->part_axis = -1;
  for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
  c->ipc_opened.clear(); c->peer_lower = nullptr; c->peer_upper = nullptr;
  if (c->d_halo_recv) { cudaFree(c->d_halo_recv); c->d_halo_recv = nullptr; }
  halo_lists_free(c);
  c->finalized = false; c->ndof = 0; c->ncon = 0;
}

__global__ void k_halfspace(MeshDesc m, int axis, double threshold, int sd_below, int sd_above, uint8_t* sd)
{
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= m.ncells()) return;
  int ijk[3]; int64_t r = c;
  ijk[0] = r % m.n[0]; r /= m.n[0]; ijk[1] = r % m.n[1]; ijk[2] = r / m.n[1];
  double lo = m.coord(axis, ijk[axis]), up = m.coord(axis, ijk[axis] + 1);
  double centre = lo + 0.5 * (up - lo);
  sd[c] = (uint8_t)(1u << (centre > threshold ? sd_above : sd_below));
}

} // namespace mdb

using namespace mdb;

// every entry point is an NVTX range (named after the C function), every phase a nested range: timelines of nsys / ncu show the
// reference's call structure (jacobian -> volume | coupling | dirichlet rows, solve -> ...)
struct NvtxScope { explicit NvtxScope(const char* name) { nvtxRangePushA(name); } ~NvtxScope() { nvtxRangePop(); } };
#define MDB_API_BEGIN(ctx)                                                       \
  if (!(ctx)) return MDB_EINVAL;                                                 \
  try {                                                                          \
    NvtxScope nvtx_scope__(__func__);                                            \
    cudaSetDevice((ctx)->device);
#define MDB_API_END(ctx)                                                         \
    return MDB_OK;                                                               \
  } catch (mdb::Error& e) { (ctx)->err = e.what(); return e.code; }              \
  catch (std::exception& e) { (ctx)->err = e.what(); return MDB_EINVAL; }

struct mdb_ctx : mdb::Ctx {};

extern "C" {

const char* mdb_version(void) { return "mdb200 0.1 (sm_100a)"; }

mdb_ctx* mdb_create(int device)
{
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libmdb200 has no CPU fallback";
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return nullptr; }
  mdb_ctx* c = new mdb_ctx();
  c->device = device;
  try {
    MDB_CUDA(cudaSetDevice(device));
    {
      // the context stream carries the bandwidth-bound kernels and gets the highest priority; the auxiliary stream (irregular
      // rows + FP64-bound FD coupling in mdb_jacobian) fills whatever SM capacity the streaming kernels leave
      int lo = 0, hi = 0;
      MDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      static const bool aux_hi = getenv("MDB_AUX_PRIO") != nullptr;
      MDB_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, aux_hi ? lo : hi));
      MDB_CUDA(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, aux_hi ? hi : lo));
    }
    MDB_CUDA(cudaEventCreate(&c->ev0));
    MDB_CUDA(cudaEventCreate(&c->ev1));
    { int lo = 0, hi = 0; MDB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi)); MDB_CUDA(cudaStreamCreateWithPriority(&c->copy_stream, cudaStreamNonBlocking, hi)); }
    MDB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
    MDB_CUDA(cudaEventCreateWithFlags(&c->ev_stage_free, cudaEventDisableTiming));
    c->d_Ae = dalloc<double>(MDB_MAX_CHILDREN * MDB_MAX_SP_PER_CHILD * MDB_MAXLOC * MDB_MAXLOC);
    c->d_scal = dalloc<double>(64);
    c->d_partial = dalloc<double>(8 * 4096);
    MDB_CUDA(cudaMallocHost(&c->h_pinned, 64 * sizeof(double)));
  } catch (std::exception& ex) { g_create_error = ex.what(); delete c; return nullptr; }
  return c;
}

void mdb_destroy(mdb_ctx* c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  sring_release(c);
  nccl_comm_release(c);
  reset_spaces(c);
  um_free(c);
  dfree(c->d_cell_sd); dfree(c->d_Ae); dfree(c->d_ftab); dfree(c->d_ftab_g); dfree(c->d_dgtab); dfree(c->d_scal); dfree(c->d_fill_ctr); dfree(c->d_partial);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->d_halo_counter) cudaFree(c->d_halo_counter);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->aux_stream) { cudaStreamSynchronize(c->aux_stream); cudaStreamDestroy(c->aux_stream); }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_copy_done) cudaEventDestroy(c->ev_copy_done);
  if (c->ev_stage_free) cudaEventDestroy(c->ev_stage_free);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* mdb_last_error(mdb_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int mdb_mesh_structured(mdb_ctx* c, int dim, const int64_t* n, const double* lower, const double* upper)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(dim == 2 || dim == 3, MDB_EINVAL, "mdb_mesh_structured: dim must be 2 or 3");
  reset_spaces(c);
  um_free(c);
  c->children.clear(); c->sps.clear(); c->cps.clear(); c->ufe = false;
  MeshDesc& m = c->mesh;
  m = MeshDesc();
  m.dim = dim;
  int64_t total = 1;
  for (int d = 0; d < 3; ++d) {
    if (d < dim) {
      MDB_REQUIRE(n[d] >= 1 && n[d] < (1 << 20), MDB_EINVAL, "mdb_mesh_structured: bad cell count");
      MDB_REQUIRE(upper[d] > lower[d], MDB_EINVAL, "mdb_mesh_structured: empty box");
      m.n[d] = (int)n[d]; m.lower[d] = lower[d]; m.h[d] = (upper[d] - lower[d]) / n[d];
    } else { m.n[d] = 1; m.lower[d] = 0; m.h[d] = 1; }
    m.gn[d] = m.n[d]; m.off[d] = 0; m.glower[d] = m.lower[d];
    total *= m.n[d];
  }
  MDB_REQUIRE(total < (int64_t)2000000000, MDB_EINVAL, "mdb_mesh_structured: more than 2e9 cells per context");
  m.partitioned = 0;
  dfree(c->d_cell_sd);
  c->d_cell_sd = dalloc<uint8_t>(total);
  c->have_mesh = true; c->have_sd = false;
  MDB_API_END(c)
}

int mdb_mesh_partition(mdb_ctx* c, const int64_t* global_n, const int64_t* cell_offset,
                       const double* global_lower, const double* global_upper)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_mesh && !c->finalized, MDB_ESTATE, "mdb_mesh_partition: call after mdb_mesh_structured, before mdb_finalize");
  MDB_REQUIRE(!c->um, MDB_EINVAL, "mdb_mesh_partition: slab partitions exist for structured meshes only (unstructured meshes: single GPU in this build)");
  MeshDesc& m = c->mesh;
  for (int d = 0; d < m.dim; ++d) {
    MDB_REQUIRE(cell_offset[d] >= 0 && cell_offset[d] + m.n[d] <= global_n[d], MDB_EINVAL, "mdb_mesh_partition: slab outside the global mesh");
    m.gn[d] = (int)global_n[d]; m.off[d] = (int)cell_offset[d]; m.glower[d] = global_lower[d];
    m.h[d] = (global_upper[d] - global_lower[d]) / global_n[d];
  }
  m.partitioned = 1;
  MDB_API_END(c)
}

int mdb_mesh_unstructured(mdb_ctx* c, int dim, int64_t nv, const double* xyz, int64_t ne, int etype, const int64_t* conn)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(dim == 2 && etype == MDB_ETYPE_TRIANGLE, MDB_EINVAL, "mdb_mesh_unstructured: this build has device kernels for 2-D triangle meshes only");
  MDB_REQUIRE(xyz && conn, MDB_EINVAL, "mdb_mesh_unstructured: null pointer");
  reset_spaces(c);
  c->children.clear(); c->sps.clear(); c->cps.clear(); c->ufe = false;
  um_ingest(c, nv, xyz, ne, conn);
  MeshDesc& m = c->mesh;
  m = MeshDesc();
  m.dim = 2;
  m.n[0] = (int)ne; m.n[1] = 1; m.n[2] = 1;          // ncells() = ne: the per-cell subdomain calls are shared with the structured path
  for (int d = 0; d < 3; ++d) { m.gn[d] = m.n[d]; m.off[d] = 0; m.lower[d] = 0; m.glower[d] = 0; m.h[d] = 1; }
  m.partitioned = 0;
  dfree(c->d_cell_sd);
  c->d_cell_sd = dalloc<uint8_t>(ne);
  c->have_mesh = true; c->have_sd = false;
  MDB_API_END(c)
}

int mdb_subdomains(mdb_ctx* c, const uint8_t* bits)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_mesh && !c->finalized, MDB_ESTATE, "mdb_subdomains: need a mesh, before mdb_finalize");
  MDB_CUDA(cudaMemcpyAsync(c->d_cell_sd, bits, c->mesh.ncells(), cudaMemcpyHostToDevice, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->have_sd = true;
  MDB_API_END(c)
}

int mdb_cell_groups(mdb_ctx* c, const int32_t* groups)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_mesh && c->um && groups, MDB_ESTATE, "mdb_cell_groups: needs an unstructured mesh");
  MDB_CUDA(cudaMemcpyAsync(c->um->d_group, groups, c->um->ne * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  MDB_API_END(c)
}

int mdb_get_subdomains(mdb_ctx* c, uint8_t* bits)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_sd && bits, MDB_ESTATE, "mdb_get_subdomains: subdomains not set");
  MDB_CUDA(cudaMemcpyAsync(bits, c->d_cell_sd, c->mesh.ncells(), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->d2h_bytes += c->mesh.ncells();
  MDB_API_END(c)
}

int mdb_subdomains_halfspace(mdb_ctx* c, int axis, double threshold, int sd_below, int sd_above)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_mesh && !c->finalized, MDB_ESTATE, "mdb_subdomains_halfspace: need a mesh, before mdb_finalize");
  MDB_REQUIRE(axis >= 0 && axis < c->mesh.dim && sd_below >= 0 && sd_below < 8 && sd_above >= 0 && sd_above < 8, MDB_EINVAL,
              "mdb_subdomains_halfspace: bad axis or subdomain index");
  MDB_REQUIRE(!c->um, MDB_EINVAL, "mdb_subdomains_halfspace: structured meshes only (use mdb_subdomains)");
  MDB_LAUNCH(c, k_halfspace, cdiv(c->mesh.ncells(), 256), 256, 0, c->mesh, axis, threshold, sd_below, sd_above, c->d_cell_sd);
  c->have_sd = true;
  MDB_API_END(c)
}

int mdb_add_space(mdb_ctx* c, int fe_kind, int subdomain)
{
  if (!c) return MDB_EINVAL;
  try {
    MDB_REQUIRE(c->have_mesh && !c->finalized, MDB_ESTATE, "mdb_add_space: need a mesh, before mdb_finalize");
    const bool simplex_fe = fe_kind == MDB_FE_P1 || fe_kind == MDB_FE_P2 || (fe_kind >= MDB_FE_OPB1 && fe_kind <= MDB_FE_OPB3);
    MDB_REQUIRE(fe_kind == MDB_FE_Q1 || fe_kind == MDB_FE_Q2 || fe_kind == MDB_FE_DGQ1 || fe_kind == MDB_FE_DGQ2 || simplex_fe, MDB_EINVAL, "mdb_add_space: unknown finite element");
    MDB_REQUIRE(simplex_fe == (c->um != nullptr), MDB_EINVAL, "mdb_add_space: simplex meshes carry the Pk / OPB spaces, structured cube meshes the Qk / QkDG spaces");
    if (simplex_fe && fe_kind != MDB_FE_P1) c->ufe = true;
    MDB_REQUIRE(subdomain >= -1 && subdomain < 8, MDB_EINVAL, "mdb_add_space: bad subdomain");
    MDB_REQUIRE((int)c->children.size() < MDB_MAX_CHILDREN, MDB_EINVAL, "mdb_add_space: too many child spaces");
    Child ch; ch.fe_kind = fe_kind; ch.k = (fe_kind == MDB_FE_Q1 || fe_kind == MDB_FE_DGQ1) ? 1 : 2; ch.subdomain = subdomain;
    ch.dg = fe_kind == MDB_FE_DGQ1 || fe_kind == MDB_FE_DGQ2;
    c->children.push_back(ch);
    return (int)c->children.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_add_coupling_space(mdb_ctx* c, int fe_kind, int left_cond_kind, uint32_t left_cond_mask, int right_cond_kind, uint32_t right_cond_mask)
{
  if (!c) return MDB_EINVAL;
  try {
    MDB_REQUIRE(!c->um, MDB_EINVAL, "mdb_add_coupling_space: not supported on unstructured meshes in this build");
    MDB_REQUIRE(c->have_mesh && !c->finalized, MDB_ESTATE, "mdb_add_coupling_space: need a mesh, before mdb_finalize");
    MDB_REQUIRE(fe_kind == MDB_FE_Q1, MDB_EINVAL, "mdb_add_coupling_space: only Q1 coupling spaces (vertex-attached mortar DOFs) have device kernels");
    MDB_REQUIRE(c->mesh.dim >= 2, MDB_EINVAL, "mdb_add_coupling_space: needs dim >= 2");
    auto ok = [](int k) { return k == MDB_COND_EQUALITY || k == MDB_COND_SUBSET; };
    MDB_REQUIRE(ok(left_cond_kind) && ok(right_cond_kind), MDB_EINVAL, "mdb_add_coupling_space: bad condition");
    MDB_REQUIRE((int)c->children.size() < MDB_MAX_CHILDREN, MDB_EINVAL, "mdb_add_coupling_space: too many child spaces");
    Child ch; ch.fe_kind = fe_kind; ch.k = 1; ch.subdomain = MDB_SUBDOMAIN_IFACE; ch.iface = true;
    ch.lkind = left_cond_kind; ch.lmask = left_cond_mask; ch.rkind = right_cond_kind; ch.rmask = right_cond_mask;
    c->children.push_back(ch);
    return (int)c->children.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_add_subproblem(mdb_ctx* c, int op_id, const void* params, size_t nbytes, int cond_kind, uint32_t cond_mask,
                       const int* children, int nchildren)
{
  if (!c) return MDB_EINVAL;
  try {
    MDB_REQUIRE(!c->finalized, MDB_ESTATE, "mdb_add_subproblem: before mdb_finalize");
    MDB_REQUIRE(cond_kind == MDB_COND_EQUALITY || cond_kind == MDB_COND_SUBSET, MDB_EINVAL, "mdb_add_subproblem: bad condition");
    MDB_REQUIRE(nchildren >= 1 && nchildren <= 4, MDB_EINVAL, "mdb_add_subproblem: 1 to 4 child spaces");
    for (int k = 0; k < nchildren; ++k) {
      MDB_REQUIRE(children[k] >= 0 && children[k] < (int)c->children.size(), MDB_EINVAL, "mdb_add_subproblem: bad child index");
      MDB_REQUIRE(!c->children[children[k]].iface, MDB_EINVAL, "mdb_add_subproblem: a coupling space cannot carry a subproblem");
    }
    SubProblemDesc sp; sp.op_id = op_id; sp.cond_kind = cond_kind; sp.mask = cond_mask;
    sp.children.assign(children, children + nchildren);
    if (op_id == MDB_OP_TAYLORHOOD_STOKES || op_id == MDB_OP_DARCY_DG) {
      // the Stokes-Darcy subproblems of test/stokesdarcy2.cc (generic element path, ufe.cu); multi-child: subproblemlocalfunctionspace.hh:144-261
      MDB_REQUIRE(c->um, MDB_EINVAL, "mdb_add_subproblem: Taylor-Hood Stokes / Darcy DG need an unstructured triangle mesh");
      if (op_id == MDB_OP_TAYLORHOOD_STOKES) {
        MDB_REQUIRE(nbytes == sizeof(mdb_taylorhood_params), MDB_EINVAL, "mdb_add_subproblem: bad mdb_taylorhood_params size");
        std::memcpy(&sp.th, params, nbytes);
        MDB_REQUIRE(nchildren == 3 && c->children[children[0]].fe_kind == MDB_FE_P2 && c->children[children[1]].fe_kind == MDB_FE_P2 &&
                    c->children[children[2]].fe_kind == MDB_FE_P1, MDB_EINVAL, "mdb_add_subproblem: Taylor-Hood needs the children {P2, P2, P1}");
        MDB_REQUIRE(sp.th.quad_order >= 2 && sp.th.quad_order <= 6, MDB_EINVAL, "mdb_add_subproblem: Taylor-Hood quadrature order 2..6");
      } else {
        MDB_REQUIRE(nbytes == sizeof(mdb_darcy_params), MDB_EINVAL, "mdb_add_subproblem: bad mdb_darcy_params size");
        std::memcpy(&sp.darcy, params, nbytes);
        const int fe = c->children[children[0]].fe_kind;
        MDB_REQUIRE(nchildren == 1 && fe >= MDB_FE_OPB1 && fe <= MDB_FE_OPB3, MDB_EINVAL, "mdb_add_subproblem: the Darcy DG operator needs one MDB_FE_OPBk child");
        MDB_REQUIRE(sp.darcy.intorderadd >= 0 && sp.darcy.intorderadd + 2 * (fe - 30) <= 6, MDB_EINVAL, "mdb_add_subproblem: DG quadrature order up to 6");
        MDB_REQUIRE(sp.darcy.method == MDB_DG_METHOD_SIPG || sp.darcy.method == MDB_DG_METHOD_NIPG, MDB_EINVAL, "mdb_add_subproblem: bad DG method");
      }
      c->ufe = true;
      c->sps.push_back(sp);
      return (int)c->sps.size() - 1;
    }
    MDB_REQUIRE(nchildren == 1, MDB_EINVAL, "mdb_add_subproblem: this local operator takes one child space");
    if (op_id == MDB_OP_POISSON) {
      MDB_REQUIRE(nbytes == sizeof(mdb_poisson_params), MDB_EINVAL, "mdb_add_subproblem: bad mdb_poisson_params size");
      std::memcpy(&sp.poisson, params, nbytes);
    } else if (op_id == MDB_OP_L2) {
      MDB_REQUIRE(nbytes == sizeof(mdb_l2_params), MDB_EINVAL, "mdb_add_subproblem: bad mdb_l2_params size");
      std::memcpy(&sp.l2, params, nbytes);
    } else if (op_id == MDB_OP_CONVDIFF_DG) {
      MDB_REQUIRE(nbytes == sizeof(mdb_convdiffdg_params), MDB_EINVAL, "mdb_add_subproblem: bad mdb_convdiffdg_params size");
      std::memcpy(&sp.cddg, params, nbytes);
      MDB_REQUIRE(c->children[children[0]].dg, MDB_EINVAL, "mdb_add_subproblem: ConvectionDiffusionDG has device kernels on QkDG spaces only");
      MDB_REQUIRE(sp.cddg.method == MDB_DG_METHOD_SIPG || sp.cddg.method == MDB_DG_METHOD_NIPG, MDB_EINVAL, "mdb_add_subproblem: bad DG method");
      sp.dg_k = c->children[children[0]].k;
    } else throw Error(MDB_EINVAL, "mdb_add_subproblem: unknown local operator (closed operator set, see mdb200.h)");
    MDB_REQUIRE(c->children[children[0]].dg == (op_id == MDB_OP_CONVDIFF_DG), MDB_EINVAL,
                "mdb_add_subproblem: QkDG spaces carry ConvectionDiffusionDG, conforming spaces Poisson / L2 (closed operator set)");
    (void)gauss1d(sp.quad_order());
    if (c->um) {
      MDB_REQUIRE(op_id == MDB_OP_POISSON || op_id == MDB_OP_L2, MDB_EINVAL, "mdb_add_subproblem: unstructured meshes carry Poisson / L2 in this build");
      MDB_REQUIRE(sp.quad_order() >= 0 && sp.quad_order() <= 4, MDB_EINVAL, "mdb_add_subproblem: simplex quadrature rules up to order 4");
    }
    c->sps.push_back(sp);
    return (int)c->sps.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_add_coupling(mdb_ctx* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int jac_mode)
{
  if (!c) return MDB_EINVAL;
  try {
    MDB_REQUIRE(!c->finalized, MDB_ESTATE, "mdb_add_coupling: before mdb_finalize");
    MDB_REQUIRE(local_sp >= 0 && local_sp < (int)c->sps.size() && remote_sp >= 0 && remote_sp < (int)c->sps.size(), MDB_EINVAL,
                "mdb_add_coupling: bad subproblem index");
    MDB_REQUIRE(jac_mode == MDB_JAC_FD_BITMATCH || jac_mode == MDB_JAC_ANALYTIC, MDB_EINVAL, "mdb_add_coupling: bad jac_mode");
    CouplingDesc cp; cp.op_id = op_id; cp.local_sp = local_sp; cp.remote_sp = remote_sp; cp.jac_mode = jac_mode;
    if (op_id == MDB_OP_CVCF_COUPLING) {
      MDB_REQUIRE(nbytes == sizeof(mdb_cvcf_params), MDB_EINVAL, "mdb_add_coupling: bad mdb_cvcf_params size");
      std::memcpy(&cp.cvcf, params, nbytes);
      (void)gauss1d(cp.cvcf.quad_order);
    } else if (op_id == MDB_OP_PROPFLOW_COUPLING) {
      MDB_REQUIRE(nbytes == sizeof(mdb_propflow_params), MDB_EINVAL, "mdb_add_coupling: bad mdb_propflow_params size");
      std::memcpy(&cp.propflow, params, nbytes);
    } else if (op_id == MDB_OP_ND_COUPLING) {
      MDB_REQUIRE(nbytes == sizeof(mdb_ndcoupling_params), MDB_EINVAL, "mdb_add_coupling: bad mdb_ndcoupling_params size");
      std::memcpy(&cp.nd, params, nbytes);
      MDB_REQUIRE(cp.nd.mode >= MDB_ND_MODE_NEUMANN && cp.nd.mode <= MDB_ND_MODE_BOTH, MDB_EINVAL, "mdb_add_coupling: bad Neumann-Dirichlet coupling mode");
      MDB_REQUIRE(!c->mesh.partitioned, MDB_EINVAL, "mdb_add_coupling: the Neumann-Dirichlet coupling is not supported on a partitioned mesh");
    } else if (op_id == MDB_OP_STOKESDARCY_COUPLING) {
      MDB_REQUIRE(nbytes == sizeof(mdb_stokesdarcy_params), MDB_EINVAL, "mdb_add_coupling: bad mdb_stokesdarcy_params size");
      std::memcpy(&cp.sd, params, nbytes);
      MDB_REQUIRE(c->sps[local_sp].op_id == MDB_OP_TAYLORHOOD_STOKES && c->sps[remote_sp].op_id == MDB_OP_DARCY_DG, MDB_EINVAL,
                  "mdb_add_coupling: the Stokes-Darcy coupling links a Taylor-Hood subproblem (local) with a Darcy DG subproblem (remote)");
      MDB_REQUIRE(cp.sd.epsilon > 0.0 && cp.sd.porosity != 0.0 && cp.sd.kabs > 0.0, MDB_EINVAL, "mdb_add_coupling: bad Stokes-Darcy parameters");
      c->cps.push_back(cp);
      return (int)c->cps.size() - 1;
    } else throw Error(MDB_EINVAL, "mdb_add_coupling: unknown coupling operator");
    MDB_REQUIRE(!c->um || op_id == MDB_OP_PROPFLOW_COUPLING || op_id == MDB_OP_CVCF_COUPLING, MDB_EINVAL,
                "mdb_add_coupling: unstructured meshes carry the two proportional-flow couplings in this build");
    MDB_REQUIRE(!c->um || op_id != MDB_OP_CVCF_COUPLING || (cp.cvcf.quad_order >= 0 && cp.cvcf.quad_order <= 5), MDB_EINVAL,
                "mdb_add_coupling: face rules up to order 5 on simplex meshes");
    c->cps.push_back(cp);
    return (int)c->cps.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_add_enriched_coupling(mdb_ctx* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int coupling_child, int jac_mode)
{
  if (!c) return MDB_EINVAL;
  try {
    MDB_REQUIRE(!c->finalized, MDB_ESTATE, "mdb_add_enriched_coupling: before mdb_finalize");
    MDB_REQUIRE(local_sp >= 0 && local_sp < (int)c->sps.size() && remote_sp >= 0 && remote_sp < (int)c->sps.size(), MDB_EINVAL,
                "mdb_add_enriched_coupling: bad subproblem index");
    MDB_REQUIRE(coupling_child >= 0 && coupling_child < (int)c->children.size() && c->children[coupling_child].iface, MDB_EINVAL,
                "mdb_add_enriched_coupling: coupling_child must come from mdb_add_coupling_space");
    MDB_REQUIRE(jac_mode == MDB_JAC_FD_BITMATCH || jac_mode == MDB_JAC_ANALYTIC, MDB_EINVAL, "mdb_add_enriched_coupling: bad jac_mode");
    MDB_REQUIRE(op_id == MDB_OP_MORTAR_POISSON_COUPLING, MDB_EINVAL, "mdb_add_enriched_coupling: unknown enriched coupling operator");
    MDB_REQUIRE(nbytes == sizeof(mdb_mortar_params), MDB_EINVAL, "mdb_add_enriched_coupling: bad mdb_mortar_params size");
    CouplingDesc cp; cp.op_id = op_id; cp.local_sp = local_sp; cp.remote_sp = remote_sp; cp.jac_mode = jac_mode; cp.coupling_child = coupling_child;
    std::memcpy(&cp.mortar, params, nbytes);
    c->cps.push_back(cp);
    return (int)c->cps.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_add_power_coupling_space(mdb_ctx* c, int fe_kind, int left_cond_kind, uint32_t left_cond_mask, int right_cond_kind, uint32_t right_cond_mask,
                                 int ncomp)
{
  if (!c) return MDB_EINVAL;
  if (ncomp < 1 || ncomp > 2) { c->err = "mdb_add_power_coupling_space: 1 or 2 component blocks"; return MDB_EINVAL; }
  int first = -1;
  for (int k = 0; k < ncomp; ++k) {
    const int ch = mdb_add_coupling_space(c, fe_kind, left_cond_kind, left_cond_mask, right_cond_kind, right_cond_mask);
    if (ch < 0) return ch;
    if (k == 0) first = ch;
  }
  return first;
}

int mdb_add_power_enriched_coupling(mdb_ctx* c, int op_id, const void* params, size_t nbytes, int local_sp, int remote_sp, int first_coupling_child,
                                    int ncomp, int jac_mode)
{
  if (!c) return MDB_EINVAL;
  if (ncomp < 1 || ncomp > 2 || first_coupling_child < 0 || first_coupling_child + ncomp > (int)c->children.size()) {
    c->err = "mdb_add_power_enriched_coupling: bad power coupling space"; return MDB_EINVAL;
  }
  for (int k = 0; k < ncomp; ++k)
    if (!c->children[first_coupling_child + k].iface) { c->err = "mdb_add_power_enriched_coupling: bad power coupling space"; return MDB_EINVAL; }
  const int h = mdb_add_enriched_coupling(c, op_id, params, nbytes, local_sp, remote_sp, first_coupling_child, jac_mode);
  if (h >= 0) c->cps[h].ncomp = ncomp;
  return h;
}

int mdb_finalize(mdb_ctx* c, int64_t* ndof)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->have_mesh && c->have_sd, MDB_ESTATE, "mdb_finalize: mesh and subdomains must be set");
  MDB_REQUIRE(!c->children.empty(), MDB_ESTATE, "mdb_finalize: no child spaces");
  c->phase_begin("dofmap");
  if (c->um && c->ufe) {
    for (const SubProblemDesc& sp : c->sps)
      MDB_REQUIRE(sp.op_id == MDB_OP_TAYLORHOOD_STOKES || sp.op_id == MDB_OP_DARCY_DG, MDB_EINVAL,
                  "mdb_finalize: P2 / OPB spaces carry the Taylor-Hood and Darcy DG subproblems (closed operator set)");
    for (const CouplingDesc& cp : c->cps) MDB_REQUIRE(cp.op_id == MDB_OP_STOKESDARCY_COUPLING, MDB_EINVAL, "mdb_finalize: P2 / OPB spaces couple through MDB_OP_STOKESDARCY_COUPLING");
    ufe_build_dofmap(c);
  } else if (c->um) um_build_dofmap(c); else build_dofmap(c);
  c->phase_end();
  if (c->mesh.partitioned) {
    // the add_* calls refuse these on a partitioned mesh; mdb_mesh_partition may also come AFTER them, so check again here.
    // Conforming Qk blocks: ownership by lattice planes and plane halos (setup_partition).  Cell-blocked QkDG blocks need the face
    // neighbours of every owned cell, i.e. ghost cell layers on both sides of the slab: mdb_mesh_partition then only declares the cut
    // (rank-independent coordinates, no cell sweeps over ghost cells) and ownership + halos come from mdb_partition_lists.
    bool any_dg = false, any_conforming = false;
    for (const Child& ch : c->children) (ch.dg ? any_dg : any_conforming) = true;      // coupling (mortar) blocks are vertex-attached: conforming
    MDB_REQUIRE(!(any_dg && any_conforming), MDB_EINVAL, "mdb_finalize: a partitioned mesh carries either conforming Qk or QkDG child spaces, not both");
    for (const CouplingDesc& cp : c->cps)
      MDB_REQUIRE(cp.op_id != MDB_OP_ND_COUPLING, MDB_EINVAL, "mdb_finalize: the Neumann-Dirichlet coupling is not supported on a partitioned mesh");
    if (!any_dg) setup_partition(c);
  }
  c->finalized = true;
  if (ndof) *ndof = c->ndof;
  MDB_API_END(c)
}

int mdb_get_dofmap(mdb_ctx* c, int64_t* cell_ptr, int64_t* cell_dofs)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_get_dofmap: not finalized");
  if (c->um && c->ufe) ufe_export_dofmap(c, cell_ptr, cell_dofs); else if (c->um) um_export_dofmap(c, cell_ptr, cell_dofs); else export_dofmap(c, cell_ptr, cell_dofs);
  MDB_API_END(c)
}

int mdb_get_child_offsets(mdb_ctx* c, int64_t* offsets)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_get_child_offsets: not finalized");
  for (size_t i = 0; i < c->children.size(); ++i) offsets[i] = c->children[i].offset;
  offsets[c->children.size()] = c->ndof;
  MDB_API_END(c)
}

int mdb_constraints_assemble(mdb_ctx* c, const int* subproblems, const mdb_bctype* bcs, int n, int64_t* nconstrained)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_constraints_assemble: not finalized");
  for (int i = 0; i < n; ++i) MDB_REQUIRE(subproblems[i] >= 0 && subproblems[i] < (int)c->sps.size(), MDB_EINVAL, "bad subproblem index");
  c->con_sps.assign(subproblems, subproblems + n); c->con_bcs.assign(bcs, bcs + n); c->con_recorded = true;
  c->phase_begin("constraints");
  if (c->um && c->ufe) ufe_assemble_constraints(c, subproblems, bcs, n); else if (c->um) um_assemble_constraints(c, subproblems, bcs, n); else assemble_constraints(c, subproblems, bcs, n);
  rebuild_conlist(c);
  c->phase_end();
  if (nconstrained) *nconstrained = c->ncon;
  MDB_API_END(c)
}

int mdb_set_constraints(mdb_ctx* c, int64_t n, const int64_t* dofs)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_set_constraints: not finalized");
  std::vector<uint8_t> f(c->ndof, 0);
  for (int64_t i = 0; i < n; ++i) { MDB_REQUIRE(dofs[i] >= 0 && dofs[i] < c->ndof, MDB_EINVAL, "constraint index out of range"); f[dofs[i]] = 1; }
  MDB_CUDA(cudaMemcpyAsync(c->d_constrained, f.data(), c->ndof, cudaMemcpyHostToDevice, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->con_recorded = n == 0; c->con_sps.clear(); c->con_bcs.clear();
  rebuild_conlist(c);
  MDB_API_END(c)
}

int mdb_get_constraints(mdb_ctx* c, uint8_t* flags)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_get_constraints: not finalized");
  MDB_CUDA(cudaMemcpyAsync(flags, c->d_constrained, c->ndof, cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  MDB_API_END(c)
}

int mdb_interpolate(mdb_ctx* c, const int* subproblems, const mdb_function* fns, int n, double time, double* x)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_interpolate: not finalized");
  MDB_REQUIRE(!c->ufe, MDB_EINVAL, "mdb_interpolate: no device interpolation for P2 / OPB spaces in this build (test/stokesdarcy2.cc starts from u = 0)");
  VecInOut vx(c, x, c->ndof, 0, true);
  for (int i = 0; i < n; ++i) MDB_REQUIRE(subproblems[i] >= 0 && subproblems[i] < (int)c->sps.size(), MDB_EINVAL, "bad subproblem index");
  if (c->um) um_interpolate(c, subproblems, fns, n, time, vx.d); else interpolate(c, subproblems, fns, n, time, vx.d);
  vx.commit();
  MDB_API_END(c)
}

int mdb_set_time(mdb_ctx* c, double t)
{
  if (!c) return MDB_EINVAL;
  c->time = t;
  return MDB_OK;
}

int mdb_gridoperator_create(mdb_ctx* c, const int* subproblems, int nsp, const int* couplings, int ncp)
{
  if (!c) return MDB_EINVAL;
  try {
    cudaSetDevice(c->device);
    MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_gridoperator_create: not finalized");
    GridOp go;
    for (int i = 0; i < nsp; ++i) { MDB_REQUIRE(subproblems[i] >= 0 && subproblems[i] < (int)c->sps.size(), MDB_EINVAL, "bad subproblem index"); go.sps.push_back(subproblems[i]); }
    for (int i = 0; i < ncp; ++i) { MDB_REQUIRE(couplings[i] >= 0 && couplings[i] < (int)c->cps.size(), MDB_EINVAL, "bad coupling index"); go.cps.push_back(couplings[i]); }
    if (c->um && c->ufe) ufe_build_facelists(c, go); else if (c->um) um_build_facelists(c, go); else build_facelists(c, go);
    c->gos.push_back(go);
    return (int)c->gos.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

int mdb_matrix_create(mdb_ctx* c, const int* gos, int ngos, int64_t* nnz)
{
  if (!c) return MDB_EINVAL;
  try {
    cudaSetDevice(c->device);
    MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_matrix_create: not finalized");
    Matrix m;
    for (int i = 0; i < ngos; ++i) { MDB_REQUIRE(gos[i] >= 0 && gos[i] < (int)c->gos.size(), MDB_EINVAL, "bad grid operator handle"); m.gos.push_back(gos[i]); }
    c->phase_begin("pattern");
    if (c->um && c->ufe) ufe_build_pattern(c, m); else if (c->um) um_build_pattern(c, m); else build_pattern(c, m);
    c->phase_end();
    if (nnz) *nnz = m.nnz;
    c->mats.push_back(m);
    return (int)c->mats.size() - 1;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
}

#define MDB_MAT(c, mat) MDB_REQUIRE((mat) >= 0 && (mat) < (int)(c)->mats.size(), MDB_EINVAL, "bad matrix handle"); Matrix& M = (c)->mats[(mat)];
#define MDB_GO(c, go) MDB_REQUIRE((go) >= 0 && (go) < (int)(c)->gos.size(), MDB_EINVAL, "bad grid operator handle"); GridOp& G = (c)->gos[(go)];

__global__ void k_widen_i32(const int32_t* in, int64_t* out, int64_t n)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

int mdb_matrix_get_pattern(mdb_ctx* c, int mat, int64_t* rowptr, int64_t* colidx)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  MDB_CUDA(cudaMemcpyAsync(rowptr, M.d_rowptr, (c->ndof + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  // widen the int32 device columns in chunks through scratch
  const int64_t chunk = 1 << 24;
  int64_t* d_tmp = (int64_t*)c->scratch(chunk * sizeof(int64_t));
  for (int64_t o = 0; o < M.nnz; o += chunk) {
    int64_t n = (M.nnz - o < chunk) ? M.nnz - o : chunk;
    MDB_LAUNCH(c, k_widen_i32, cdiv(n, 256), 256, 0, M.d_colidx + o, d_tmp, n);
    MDB_CUDA(cudaMemcpyAsync(colidx + o, d_tmp, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
  }
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  MDB_API_END(c)
}

int mdb_matrix_get_values(mdb_ctx* c, int mat, double* values)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  MDB_CUDA(cudaMemcpyAsync(values, M.d_val, M.nnz * sizeof(double), cudaMemcpyDefault, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  MDB_API_END(c)
}

int mdb_matrix_zero(mdb_ctx* c, int mat)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  MDB_CUDA(cudaMemsetAsync(M.d_val, 0, M.nnz * sizeof(double), c->stream));
  M.dinv_valid = false; M.ilu_valid = false; direct_invalidate(M); M.asm_log.clear();
  MDB_API_END(c)
}

int mdb_mg_setup(mdb_ctx* c, int mat, int* nlevels)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  const int n = mg_setup(c, M);
  if (nlevels) *nlevels = n;
  MDB_API_END(c)
}

mdb_ctx* mdb_mg_level(mdb_ctx* c, int mat, int level)
{
  if (!c || mat < 0 || mat >= (int)c->mats.size()) return nullptr;
  return mg_level_ctx(c->mats[mat], level);
}

int mdb_matrix_mg_levels(mdb_ctx* c, int mat)
{
  if (!c || mat < 0 || mat >= (int)c->mats.size()) return MDB_EINVAL;
  return mg_levels(c->mats[mat]);
}

int mdb_matrix_device_ptrs(mdb_ctx* c, int mat, const int64_t** rowptr, const int32_t** colidx, double** values)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  if (rowptr) *rowptr = M.d_rowptr;
  if (colidx) *colidx = M.d_colidx;
  if (values) *values = M.d_val;
  MDB_API_END(c)
}

int mdb_residual(mdb_ctx* c, int go, double weight, const double* x, double* r)
{
  MDB_API_BEGIN(c)
  MDB_GO(c, go)
  VecIn vx(c, x, c->ndof, 0);
  VecInOut vr(c, r, c->ndof, 1, true);
  if (c->um) { if (c->ufe) ufe_residual(c, G, weight, vx.d, vr.d); else um_residual(c, G, weight, vx.d, vr.d); vr.commit(); return MDB_OK; }
  static const bool serial = getenv("MDB_RESIDUAL_SERIAL") != nullptr;          // A/B switch for profiling only
  if (!serial && residual_can_split(c, G)) {
    // main (high priority): element matrices -> [fork] -> row lists (latency bound) ---------------------------------> [join] -> constrain
    // aux  (low priority) :                     [fork] -> streamed uniform rows of every child (HBM / issue bound) ---> [join]
    // third (high)        :                     [fork] -> boundary terms -> coupling terms (latency bound) -----------> [join]
    // The streamed rows are interior rows all of whose cells carry the subproblem: no row list and no boundary term touches them;
    // the coupling terms of the interface cells do (their vertices one layer off the interface), which is why the stencil kernel
    // adds with a reduction instead of load + store.  The streaming kernel fills every SM; on the LOW-priority stream it yields
    // block slots to the short latency-bound kernels of the other two branches as its blocks retire (on the high-priority stream
    // it ran first and the branches serialised behind it: 0.265 ms per call against 0.2 ms).
    residual_prepare(c, G);
    // MDB_RESIDUAL_TRACE=1: a timeline of the call's branches (events on the three streams, printed to stderr; synchronises)
    static const bool trace = getenv("MDB_RESIDUAL_TRACE") != nullptr;
    static cudaEvent_t tev[8] = {nullptr};
    if (trace && !tev[0]) for (auto& e : tev) MDB_CUDA(cudaEventCreate(&e));
    auto mark = [&](int i) { if (trace) MDB_CUDA(cudaEventRecord(tev[i], c->stream)); };
    c->phase_begin("residual_volume");
    mark(0);
    launch_element_matrices(c, G.sps, weight);
    mark(1);
    MDB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    {
      StreamScope third(c, c->copy_stream);                                     // every update of r on the three branches is a reduction
      MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
      residual_boundary(c, G, weight, vr.d);
      mark(2);
      residual_coupling(c, G, weight, vx.d, vr.d);
      mark(3);
      MDB_CUDA(cudaEventRecord(c->ev_copy_done, c->stream));
    }
    {
      StreamScope aux(c, c->aux_stream);
      MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
      c->phase_end();
      c->phase_begin("residual_streamed");                                      // events on the auxiliary stream: the streaming kernel's own time inside the call
      mark(4);
      residual_volume_part(c, G, weight, vx.d, vr.d, 0);
      mark(5);
      c->phase_end();
      MDB_CUDA(cudaEventRecord(c->ev_join, c->stream));
    }
    c->phase_begin("residual_rows");
    residual_volume_part(c, G, weight, vx.d, vr.d, 1);
    mark(6);
    c->phase_end();
    MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
    c->phase_begin("residual_coupling");
    constrain_residual(c, vr.d);
    c->phase_end();
    mark(7);
    if (trace) {
      MDB_CUDA(cudaEventSynchronize(tev[7]));
      float t[8];
      for (int i = 1; i < 8; ++i) MDB_CUDA(cudaEventElapsedTime(&t[i], tev[0], tev[i]));
      fprintf(stderr, "residual trace (us from the start): elem %.1f | third: boundary %.1f coupling %.1f | aux: streamed %.1f .. %.1f | main: rows %.1f | end %.1f\n",
              t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3, t[6] * 1e3, t[7] * 1e3);
    }
  } else {
    c->phase_begin("residual_volume");
    residual_volume(c, G, weight, vx.d, vr.d);
    c->phase_end();
    c->phase_begin("residual_boundary");
    residual_boundary(c, G, weight, vr.d);
    c->phase_end();
    c->phase_begin("residual_coupling");
    residual_coupling(c, G, weight, vx.d, vr.d);
    constrain_residual(c, vr.d);
    c->phase_end();
  }
  vr.commit();
  MDB_API_END(c)
}

// x_host[list[i]] -> stage[list[i]]: the kernel reads pinned host memory directly (zero-copy over PCIe), so only the
// coefficients the Jacobian kernels actually read cross the bus
__global__ void k_gather_needed(const double* __restrict__ x_host, const int32_t* __restrict__ list, int64_t n, double* stage)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = list[i];
    stage[g] = x_host[g];
  }
}

int mdb_jacobian(mdb_ctx* c, int go, int mat, double weight, const double* x, int overwrite)
{
  MDB_API_BEGIN(c)
  MDB_GO(c, go)
  MDB_MAT(c, mat)
  M.dinv_valid = false; M.ilu_valid = false; direct_invalidate(M);
  if (overwrite) M.asm_log.clear();
  M.asm_log.push_back(std::make_pair(go, weight));
  if (c->um) { VecIn vx(c, x, c->ndof, 0); if (c->ufe) ufe_jacobian(c, G, M, weight, vx.d, overwrite != 0); else um_jacobian(c, G, M, weight, vx.d, overwrite != 0); return MDB_OK; }
  // The volume operators of the closed set are linear: their Jacobian does not read x.  Only the finite-difference
  // coupling Jacobian does, and only at the DOFs of the interface cells.  With a HOST vector the upload therefore runs on
  // the copy stream concurrently with the volume kernels, and from pinned memory only the needed entries are fetched.
  const double* dx = x;
  const bool host_x = !is_device_ptr(x);
  const bool need_x = !G.cps.empty();
  if (host_x && need_x) {
    if (!c->d_stage[0]) c->d_stage[0] = dalloc<double>(c->ndof);
    cudaPointerAttributes a;
    bool pinned = cudaPointerGetAttributes(&a, x) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
    if (!pinned) cudaGetLastError();
    static const bool full_upload = getenv("MDB_FULL_UPLOAD") != nullptr;     // measurement switch: always copy the whole vector
    if (full_upload) pinned = false;
    MDB_CUDA(cudaEventRecord(c->ev_stage_free, c->stream));                   // earlier calls may still read the staging vector
    MDB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_stage_free, 0));
    if (pinned) {
      if (G.n_xneed < 0) build_xneed(c, G);
      if (G.n_xneed > 0) {
        k_gather_needed<<<cdiv(G.n_xneed, 256) < 592 ? cdiv(G.n_xneed, 256) : 592, 256, 0, c->copy_stream>>>((const double*)a.devicePointer, G.d_xneed,
                                                                                                       G.n_xneed, c->d_stage[0]);
        c->launches++;
        MDB_CUDA(cudaGetLastError());
      }
      c->h2d_bytes += 8 * G.n_xneed;
    } else {
      MDB_CUDA(cudaMemcpyAsync(c->d_stage[0], x, c->ndof * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
      c->h2d_bytes += 8 * c->ndof;
    }
    MDB_CUDA(cudaEventRecord(c->ev_copy_done, c->copy_stream));
    dx = c->d_stage[0];
  }
  // Stream plan: main  : element matrices -> [fork] -> simple rows (HBM-write bound) ------------------> [join] -> Dirichlet rows
  //              aux   :                      [fork] -> irregular rows -> (x ready) -> FD coupling (FP64 bound) -> [join]
  // The two branches write disjoint rows (coupling entries live in irregular rows only) and stress different units, so
  // they overlap almost perfectly.
  c->phase_begin("jacobian_elem");
  launch_element_matrices(c, G.sps, weight);
  c->phase_end();
  static const bool rows_first = getenv("MDB_JAC_ROWS_FIRST") != nullptr;   // A/B: irregular rows alone before the fork
  if (rows_first) {
    c->phase_begin("jacobian_rows");
    jacobian_volume(c, G, M, weight, overwrite != 0, 1);
    c->phase_end();
  }
  static const bool serial_env = getenv("MDB_JAC_SERIAL") != nullptr;       // A/B: the auxiliary branch on the main stream (stand-alone kernel times)
  static const bool overlap_env = getenv("MDB_JAC_OVERLAP") != nullptr;     // A/B: two streams whatever the couplings
  // Enriched (mortar) couplings: their FD kernel is bound by the latency of serial FP64 chains at 12 warps per SM; beside the fill both
  // get slower than one after the other (256^3: fill 0.64 -> 0.90 ms, coupling 0.90 -> 1.57 ms, step 1.79 against 1.67 ms serial,
  // profiles/r02/mortar_r4.md), so such a step runs on one stream.
  bool serial = serial_env;
  if (!overlap_env) for (int q : G.cps) serial = serial || c->cps[q].enriched();
  // Dirichlet rows beside the fill (auxiliary stream) when every writer of a constrained row is on that stream: no late coupling
  // (Neumann-Dirichlet) and no constrained row among the streamed ones
  static const bool dir_late = getenv("MDB_JAC_DIRICHLET_LATE") != nullptr;  // A/B: the pass after the join, as in round 1
  bool dir_early = !dir_late && !serial && !rows_first;
  for (int q : G.cps) dir_early = dir_early && c->cps[q].op_id != MDB_OP_ND_COUPLING;
  for (const Child& ch : c->children) dir_early = dir_early && ch.fe_kind == MDB_FE_Q1;
  dir_early = dir_early && dirichlet_rows_off_the_fill(c, M);
  MDB_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  {
    StreamScope aux(c, serial ? c->stream : c->aux_stream);
    MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
    if (!rows_first) {
      c->phase_begin("jacobian_rows");
      jacobian_volume(c, G, M, weight, overwrite != 0, 1);
      c->phase_end();
    }
    if (host_x && need_x) MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
    c->phase_begin("jacobian_coupling");
    jacobian_coupling(c, G, M, weight, dx, false);
    c->phase_end();
    if (dir_early) {
      c->phase_begin("jacobian_dirichlet");
      dirichlet_rows(c, M);
      c->phase_end();
    }
    MDB_CUDA(cudaEventRecord(c->ev_join, c->stream));
  }
  c->phase_begin("jacobian_fill");
  jacobian_volume(c, G, M, weight, overwrite != 0, 0);
  c->phase_end();
  MDB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  jacobian_coupling(c, G, M, weight, dx, true);            // couplings without a pattern of their own: after every volume row is written
  if (!dir_early) {
    c->phase_begin("jacobian_dirichlet");
    dirichlet_rows(c, M);
    c->phase_end();
  }
  // a host x is only read during the call (mdb200.h): the gather / upload from page-locked memory is asynchronous, so wait for it
  // (the copy stream's work is a few microseconds and long finished; the assembly kernels keep running)
  if (host_x && need_x) MDB_CUDA(cudaEventSynchronize(c->ev_copy_done));
  MDB_API_END(c)
}

int mdb_matrix_stats(mdb_ctx* c, int mat, int64_t* simple_rows, int64_t* irregular_rows)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  int64_t irr = 0;
  for (size_t i = 0; i < c->children.size(); ++i) irr += M.n_irregular[i];
  if (irregular_rows) *irregular_rows = irr;
  if (simple_rows) *simple_rows = c->ndof - irr;
  MDB_API_END(c)
}

int mdb_bytes_copied(mdb_ctx* c, int64_t* h2d, int64_t* d2h, int reset)
{
  if (!c) return MDB_EINVAL;
  if (h2d) *h2d = c->h2d_bytes;
  if (d2h) *d2h = c->d2h_bytes;
  if (reset) { c->h2d_bytes = 0; c->d2h_bytes = 0; }
  return MDB_OK;
}

int mdb_spmv(mdb_ctx* c, int mat, const double* x, double* y)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  VecIn vx(c, x, c->ndof, 0);
  VecInOut vy(c, y, c->ndof, 1, false);
  c->phase_begin("spmv");
  if (c->halo_lists && c->nranks > 1) {
    MDB_REQUIRE(is_device_ptr(x), MDB_EINVAL, "mdb_spmv: x must be a device vector on a partitioned context");
    halo_exchange(c, const_cast<double*>(vx.d));
    spmv(c, M, vx.d, vy.d);
  } else if (c->peer_lower || c->peer_upper) {
    MDB_REQUIRE(is_device_ptr(x), MDB_EINVAL, "mdb_spmv: x must be a device vector on a partitioned context");
    spmv_partitioned(c, M, const_cast<double*>(vx.d), vy.d);      // push -> interior rows -> wait -> boundary rows
  } else spmv(c, M, vx.d, vy.d);
  c->phase_end();
  vy.commit();
  MDB_API_END(c)
}

int mdb_solve(mdb_ctx* c, int mat, int method, int precond, double reduction, int maxit, const double* rhs, double* z,
              mdb_solve_stats* stats)
{
  if (!c) return MDB_EINVAL;
  try {
    cudaSetDevice(c->device);
    MDB_MAT(c, mat)
    VecIn vb(c, rhs, c->ndof, 0);
    VecInOut vz(c, z, c->ndof, 1, false);
    int rc = solve(c, M, method, precond, reduction, maxit, -1, vb.d, vz.d, stats);
    vz.commit();
    return rc;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
  catch (std::exception& e) { c->err = e.what(); return MDB_EINVAL; }
}

// flag[0] = 1 if the matrix has an entry whose row and column lie on different sides of the block boundary
__global__ void k_block_decoupled(int64_t n, int64_t lo, int64_t hi, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int* flag)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const bool rin = r >= lo && r < hi;
  for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
    const int32_t cidx = colidx[p];
    if ((cidx >= lo && cidx < hi) != rin) { *flag = 1; return; }
  }
}
__global__ void k_copy_block(const double* __restrict__ src, double* dst, int64_t n, int64_t lo, int64_t hi, int zero_outside)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i >= lo && i < hi) dst[i] = src[i];
  else if (zero_outside) dst[i] = 0.0;
}

int mdb_solve_block(mdb_ctx* c, int mat, int child, int method, int precond, double reduction, int maxit, const double* rhs, double* z,
                    mdb_solve_stats* stats)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  MDB_REQUIRE(child >= 0 && child < (int)c->children.size(), MDB_EINVAL, "mdb_solve_block: bad child index");
  MDB_REQUIRE(c->nranks == 1, MDB_EINVAL, "mdb_solve_block: single-rank contexts only");
  const int64_t n = c->ndof, lo = c->children[child].offset, hi = lo + c->children[child].size;
  // The Krylov loop runs on the whole vector with the right-hand side zeroed outside the block.  That IS the solve of the
  // diagonal block provided the matrix does not couple the block with the rest (then every iterate stays exactly zero outside
  // the block and every scalar of the iteration is the block's own); rows without entries act as identity rows.
  int* d_flag = dalloc<int>(1);
  MDB_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), c->stream));
  MDB_LAUNCH(c, k_block_decoupled, cdiv(n, 256), 256, 0, n, lo, hi, M.d_rowptr, M.d_colidx, d_flag);
  int coupled = 0;
  MDB_CUDA(cudaMemcpyAsync(&coupled, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  dfree(d_flag);
  MDB_REQUIRE(!coupled, MDB_EINVAL, "mdb_solve_block: the matrix couples the block with other blocks (assemble the block's own operator into its own matrix)");
  VecIn vr(c, rhs, n, 0);
  VecInOut vz(c, z, n, 1, true);
  if (!c->d_stage[2]) c->d_stage[2] = dalloc<double>(n);
  if (!c->d_stage[3]) c->d_stage[3] = dalloc<double>(n);
  double* br = c->d_stage[2]; double* bz = c->d_stage[3];
  MDB_LAUNCH(c, k_copy_block, cdiv(n, 256), 256, 0, vr.d, br, n, lo, hi, 1);
  MDB_CUDA(cudaMemsetAsync(bz, 0, n * sizeof(double), c->stream));
  int rc = solve(c, M, method, precond, reduction, maxit, -1, br, bz, stats);
  MDB_LAUNCH(c, k_copy_block, cdiv(n, 256), 256, 0, bz, vz.d, n, lo, hi, 0);
  vz.commit();
  if (rc != MDB_OK) return rc;
  MDB_API_END(c)
}

int mdb_set_ssor_sweeps(mdb_ctx* c, int sweeps)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(sweeps >= 1, MDB_EINVAL, "SSOR needs at least one sweep");
  c->ssor_sweeps = sweeps;
  MDB_API_END(c)
}

int mdb_solve_fixed(mdb_ctx* c, int mat, int method, int precond, int iters, const double* rhs, double* z, mdb_solve_stats* stats)
{
  if (!c) return MDB_EINVAL;
  try {
    cudaSetDevice(c->device);
    MDB_MAT(c, mat)
    VecIn vb(c, rhs, c->ndof, 0);
    VecInOut vz(c, z, c->ndof, 1, false);
    int rc = solve(c, M, method, precond, 0.0, iters, iters, vb.d, vz.d, stats);
    vz.commit();
    return rc;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
  catch (std::exception& e) { c->err = e.what(); return MDB_EINVAL; }
}

__global__ void k_copy_nonconstrained(const double* xold, double* xnew, const uint8_t* con, int64_t n)
{
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && !con[i]) xnew[i] = xold[i];
}

int mdb_copy_nonconstrained(mdb_ctx* c, const double* xold, double* xnew)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "not finalized");
  VecIn vo(c, xold, c->ndof, 0);
  VecInOut vn(c, xnew, c->ndof, 1, true);
  MDB_LAUNCH(c, k_copy_nonconstrained, cdiv(c->ndof, 256), 256, 0, vo.d, vn.d, c->d_constrained, c->ndof);
  vn.commit();
  MDB_API_END(c)
}

__global__ void k_sub(const double* __restrict__ a, const double* __restrict__ b, double* out, int64_t n)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i] - b[i];
}

// One time step the way [UPSTREAM] PDELab::OneStepMethod drives OneStepGridOperator(go0, go1)
// (test/testinstationarypoisson.cc:320-403) for the diagonally implicit schemes of [UPSTREAM]
// TimeSteppingParameterInterface: stage r = 1..s solves
//   sum_{i<r} [a_ri r1(x_i) + b_ri dt r0(x_i)]  +  a_rr r1(x) + b_rr dt r0(x) = 0,   r0 at time t + d_i dt,
// with one Newton step (the operators are linear; StationaryLinearProblemSolver): start value = new Dirichlet values by
// interpolation at the stage time + copy_nonconstrained_dofs from the previous stage (gridoperator.hh:139-153),
// J = a_rr J1 + b_rr dt J0 (setWeight, localassembler.hh:214-255), x_r = x - J^-1 R(x).  Everything stays on the device.
int mdb_onestep_apply(mdb_ctx* c, int go0, int go1, int mat, const int* subproblems, const mdb_function* g, int n, double time, double dt, int scheme,
                      double theta, int method, int precond, double reduction, int maxit, const double* u_old, double* u_new, mdb_solve_stats* stats)
{
  MDB_API_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_onestep_apply: not finalized");
  MDB_REQUIRE(go0 >= 0 && go0 < (int)c->gos.size() && go1 >= 0 && go1 < (int)c->gos.size(), MDB_EINVAL, "bad grid operator handle");
  // Butcher-like tables ([UPSTREAM] ImplicitEulerParameter, OneStepThetaParameter, Alexander2Parameter)
  int s = 1;
  double A[2][3] = {{-1.0, 1.0, 0.0}, {0.0, 0.0, 0.0}}, B[2][3] = {{0.0, 1.0, 0.0}, {0.0, 0.0, 0.0}}, D[3] = {0.0, 1.0, 0.0};
  if (scheme == MDB_ONESTEP_THETA) { B[0][0] = 1.0 - theta; B[0][1] = theta; }
  else if (scheme == MDB_ONESTEP_ALEXANDER2) {
    const double al = 1.0 - 0.5 * sqrt(2.0);
    s = 2;
    A[0][0] = -1.0; A[0][1] = 1.0; A[0][2] = 0.0; A[1][0] = -1.0; A[1][1] = 0.0; A[1][2] = 1.0;
    B[0][0] = 0.0; B[0][1] = al; B[0][2] = 0.0; B[1][0] = 0.0; B[1][1] = 1.0 - al; B[1][2] = al;
    D[0] = 0.0; D[1] = al; D[2] = 1.0;
  } else MDB_REQUIRE(scheme == MDB_ONESTEP_IMPLICIT_EULER, MDB_EINVAL, "unknown one-step scheme");
  const int64_t nd = c->ndof;
  if (c->onestep_n < nd) { dfree(c->d_onestep); c->d_onestep = dalloc<double>(4 * nd); c->onestep_n = nd; }
  double *xs[3] = {nullptr, c->d_onestep, c->d_onestep + nd}, *R = c->d_onestep + 2 * nd, *z = c->d_onestep + 3 * nd;
  VecIn vo(c, u_old, nd, 2);
  VecInOut vn(c, u_new, nd, 3, false);
  xs[0] = const_cast<double*>(vo.d);
  int rc;
  mdb_solve_stats st_total; std::memset(&st_total, 0, sizeof(st_total)); st_total.converged = 1;
  for (int r = 1; r <= s; ++r) {
    double* x = xs[r];
    const double tr = time + D[r] * dt;
    c->time = tr;
    MDB_CUDA(cudaMemsetAsync(x, 0, nd * sizeof(double), c->stream));
    if (c->um) um_interpolate(c, subproblems, g, n, tr, x); else interpolate(c, subproblems, g, n, tr, x);
    MDB_LAUNCH(c, k_copy_nonconstrained, cdiv(nd, 256), 256, 0, xs[r - 1], x, c->d_constrained, nd);
    if ((rc = mdb_jacobian(c, go1, mat, A[r - 1][r], x, 1)) != MDB_OK) return rc;
    if ((rc = mdb_jacobian(c, go0, mat, B[r - 1][r] * dt, x, 0)) != MDB_OK) return rc;
    MDB_CUDA(cudaMemsetAsync(R, 0, nd * sizeof(double), c->stream));
    for (int i = 0; i <= r; ++i) {
      const double* xi = (i == r) ? x : xs[i];
      c->time = time + D[i] * dt;
      if (A[r - 1][i] != 0.0 && (rc = mdb_residual(c, go1, A[r - 1][i], xi, R)) != MDB_OK) return rc;
      if (B[r - 1][i] != 0.0 && (rc = mdb_residual(c, go0, B[r - 1][i] * dt, xi, R)) != MDB_OK) return rc;
    }
    c->time = tr;
    mdb_solve_stats st;
    if ((rc = mdb_solve(c, mat, method, precond, reduction, maxit, R, z, &st)) != MDB_OK) return rc;
    st_total.iterations += st.iterations; st_total.converged &= st.converged; st_total.seconds += st.seconds;
    st_total.defect0 = st.defect0; st_total.defect = st.defect;
    MDB_LAUNCH(c, k_sub, 148 * 8, 256, 0, x, z, (r == s) ? vn.d : x, nd);
  }
  if (stats) *stats = st_total;
  vn.commit();
  MDB_API_END(c)
}

double mdb_phase_ms(mdb_ctx* c, const char* phase)
{
  if (!c) return -1.0;
  try { cudaSetDevice(c->device); return c->phase_mean_ms(phase, nullptr); } catch (std::exception& e) { c->err = e.what(); return -1.0; }
}

int mdb_phase_count(mdb_ctx* c, const char* phase)
{
  if (!c) return 0;
  auto it = c->phases.find(phase);
  return it == c->phases.end() ? 0 : (int)it->second.used;
}

int mdb_profile_reset(mdb_ctx* c)
{
  if (!c) return MDB_EINVAL;
  c->phase_reset();
  return MDB_OK;
}

// sum of squares of the stored matrix entries (a cheap device-side checksum of an assembly result)
// [UPSTREAM] PDELab StationaryLinearProblemSolver::apply as the reference's drivers call it (test/testpoisson.cc:295-298,
// SURVEY 8c (vii)): assemble J(x0) and r = R(x0) at the incoming x0, solve J z = r to the relative defect reduction, x <- x0 - z.
// One call, everything on the device; x may be a host vector (uploaded once, the solution is copied back) or a device vector
// (updated in place).  keep_matrix != 0 reuses the values already in `mat` (the reference's keep_matrix of the Dirichlet-Neumann loop).
int mdb_stationary_solve(mdb_ctx* c, int go, int mat, int method, int precond, double reduction, int maxit, double* x, int keep_matrix,
                         mdb_solve_stats* stats)
{
  if (!c) return MDB_EINVAL;
  try {
    cudaSetDevice(c->device);
    MDB_MAT(c, mat)
    MDB_REQUIRE(go >= 0 && go < (int)c->gos.size(), MDB_EINVAL, "bad grid operator handle");
    MDB_REQUIRE(x, MDB_EINVAL, "mdb_stationary_solve: null vector");
    const int64_t nd = c->ndof;
    if (c->onestep_n < nd) { dfree(c->d_onestep); c->d_onestep = dalloc<double>(4 * nd); c->onestep_n = nd; }
    const bool host = !is_device_ptr(x);
    double* u = host ? c->d_onestep : x;
    double* r = c->d_onestep + 2 * nd;
    double* z = c->d_onestep + 3 * nd;
    c->phase_begin("stationary_upload");
    if (host) { MDB_CUDA(cudaMemcpyAsync(u, x, nd * sizeof(double), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += nd * (int64_t)sizeof(double); }
    c->phase_end();
    int rc = MDB_OK;
    if (!keep_matrix) { rc = mdb_jacobian(c, go, mat, 1.0, u, 1); if (rc != MDB_OK) return rc; }
    MDB_CUDA(cudaMemsetAsync(r, 0, nd * sizeof(double), c->stream));
    rc = mdb_residual(c, go, 1.0, u, r);
    if (rc != MDB_OK) return rc;
    rc = solve(c, M, method, precond, reduction, maxit, -1, r, z, stats);
    if (rc != MDB_OK) return rc;
    MDB_LAUNCH(c, k_sub, 148 * 8, 256, 0, u, z, u, nd);
    if (host) {
      MDB_CUDA(cudaMemcpyAsync(x, u, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      MDB_CUDA(cudaStreamSynchronize(c->stream));
      c->d2h_bytes += nd * (int64_t)sizeof(double);
    }
    return MDB_OK;
  } catch (mdb::Error& e) { c->err = e.what(); return e.code; }
  catch (std::exception& e) { c->err = e.what(); return MDB_EINVAL; }
}

__global__ void k_norm2_partial(const double* __restrict__ v, int64_t n, double* partial)
{
  __shared__ double sh[8];
  double a = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a += v[i] * v[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i]; partial[blockIdx.x] = s; }
}
__global__ void k_norm2_final(const double* __restrict__ partial, int n, double* out)
{
  __shared__ double sh[8];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i]; out[0] = s; }
}

int mdb_matrix_norm2(mdb_ctx* c, int mat, double* out)
{
  MDB_API_BEGIN(c)
  MDB_MAT(c, mat)
  const int grid = 148 * 8;
  c->phase_begin("matrix_norm2");
  MDB_LAUNCH(c, k_norm2_partial, grid, 256, 0, M.d_val, M.nnz, c->d_partial);
  MDB_LAUNCH(c, k_norm2_final, 1, 256, 0, c->d_partial, grid, c->d_scal + 40);
  c->phase_end();
  MDB_CUDA(cudaMemcpyAsync(c->h_pinned + 40, c->d_scal + 40, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  *out = c->h_pinned[40];
  c->d2h_bytes += 8;
  MDB_API_END(c)
}

int64_t mdb_launch_count(mdb_ctx* c, int reset)
{
  if (!c) return 0;
  int64_t v = c->launches;
  if (reset) c->launches = 0;
  return v;
}

void* mdb_stream(mdb_ctx* c) { return c ? (void*)c->stream : nullptr; }

int mdb_synchronize(mdb_ctx* c)
{
  MDB_API_BEGIN(c)
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  MDB_API_END(c)
}

} // extern "C"
