// multigpu.cu — multi-GPU plumbing of the hot path: slab partition with one ghost cell layer,
// halo exchange of Krylov vectors by direct peer stores over NVLink, NCCL all-reduce of the Krylov scalars.
//
// The reference's only parallel mode is overlapping domain decomposition on an MPI grid with overlap 1
// (test/testoverlappinginstationarypoisson.cc:232-233; gridoperator.hh:65-71 "nonoverlapping_mode = false"): every rank
// assembles all cells that touch its rows, so assembly needs no communication and only vectors are exchanged
// (gridoperator.hh:150-152).  The same idea here, one process per GPU:
//   * rank r owns the cell layers [c0, c1) along the slowest axis and the vertex planes [c0, c1) (the last rank also
//     owns the top plane); its local mesh is the owned layers plus ONE ghost layer below (c0 - 1), so all cells around
//     an owned vertex are local and owned rows are assembled completely without communication;
//   * the local vector carries two ghost planes (below: owned by r-1, above: owned by r+1).  Before every SpMV the
//     owners push their boundary planes straight into the neighbours' receive windows (st.global on cudaIpc-mapped peer
//     memory, i.e. NVLink P2P stores), followed by a release-store of a sequence number; the receiver's unpack kernel
//     acquires the sequence number and copies the window into its ghost entries.  No host synchronisation, no staging;
//   * dot products are reduced over owned rows only (Ctx::d_owned) and summed with ncclAllReduce on the context stream.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already carries, e.g. torch's) so that the
// single-GPU product has no link-time dependency on it.
#include "common.cuh"
#include <dlfcn.h>
#include <unistd.h>

namespace mdb {

// ---- NCCL through dlopen ---------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid_t;
typedef void* nccl_comm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid_t*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, nccl_uid_t, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };      // ncclDouble / ncclSum of nccl.h (stable ABI values)

static NcclApi& nccl()
{
  if (g_nccl.lib) return g_nccl;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD); if (h) break; }     // the copy torch loaded
  for (const char* n : names) { if (h) break; h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
  MDB_REQUIRE(h, MDB_ENCCL, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : ""));
  g_nccl.GetUniqueId = (int (*)(nccl_uid_t*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm_t*, int, nccl_uid_t, int))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(nccl_comm_t))dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  MDB_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce, MDB_ENCCL, "libnccl lacks expected symbols");
  g_nccl.lib = h;
  return g_nccl;
}
#define MDB_NCCL(call)                                                                                                   \
  do { int r__ = (call); if (r__ != 0) throw Error(MDB_ENCCL, std::string(#call) + ": " +                                \
                                                   (nccl().GetErrorString ? nccl().GetErrorString(r__) : "nccl error")); } while (0)

// ---- peer-memory all-reduce of the Krylov scalars --------------------------------------------------------------------------------
// NCCL's all-reduce of one or two doubles costs ~14 us per call on NVSwitch (launch + protocol), two or three times per Krylov
// iteration.  Every rank instead owns a 1.3 KB window that all ranks map (cudaIpc): a block of the reducing kernel stores its values
// straight into every peer's window over NVLink and spins on its own window (sr_exchange in common.cuh) — no library call, no extra
// launch when the exchange is the tail of the kernel that produced the values (k_reduce_partials in solver.cu).
struct ScalarRing { double* mine = nullptr; double* win[SR_MAXR]; int nranks = 0, rank = 0; unsigned long long seq = 0; bool connected = false; int* d_err = nullptr; };
static const size_t SR_WINDOW_BYTES = (2 * SR_MAXR * SR_MAXV) * sizeof(double) + (2 * SR_MAXR) * sizeof(unsigned long long);

static long long halo_timeout_clocks(const Ctx* c);
bool sring_connected(const Ctx* c)
{
  static const bool off = getenv("MDB_SCALARS_NCCL") != nullptr;                 // A/B switch: always NCCL
  const ScalarRing* R = (const ScalarRing*)c->sring;
  return !off && R && R->connected && R->nranks == c->nranks;
}
SrSpec sring_next(Ctx* c, double* d_done)
{
  ScalarRing* R = (ScalarRing*)c->sring;
  SrSpec S; S.nranks = R->nranks; S.rank = R->rank; S.seq = ++R->seq;
  for (int i = 0; i < SR_MAXR; ++i) S.win[i] = i < R->nranks ? R->win[i] : nullptr;
  S.err = R->d_err; S.timeout_clocks = halo_timeout_clocks(c); S.done = d_done;
  return S;
}
void sring_release(Ctx* c)
{
  ScalarRing* R = (ScalarRing*)c->sring;
  if (R && !c->sring_borrowed) { if (R->mine) cudaFree(R->mine); if (R->d_err) cudaFree(R->d_err); delete R; }
  c->sring = nullptr; c->sring_borrowed = false;
}
__global__ void k_sr_allreduce(SrSpec R, double* vals, int n)
{
  __shared__ double loc[SR_MAXV];
  if ((int)threadIdx.x < n) loc[threadIdx.x] = vals[threadIdx.x];
  __syncthreads();
  sr_exchange(R, loc, n);
  if ((int)threadIdx.x < n) vals[threadIdx.x] = loc[threadIdx.x];
}

void allreduce_scalars(Ctx* c, double* d_vals, int n)
{
  if (sring_connected(c) && n <= SR_MAXV) {
    MDB_LAUNCH(c, k_sr_allreduce, 1, 32, 0, sring_next(c, nullptr), d_vals, n);
    return;
  }
  MDB_REQUIRE(c->nccl_comm, MDB_ENCCL, "multi-GPU solve without a communicator: call mdb_comm_init first");
  MDB_NCCL(nccl().AllReduce(d_vals, d_vals, (size_t)n, NCCL_FLOAT64, NCCL_SUM, (nccl_comm_t)c->nccl_comm, c->stream));
  c->launches++;
}

void nccl_comm_release(Ctx* c)
{
  if (c->nccl_comm && !c->comm_borrowed && g_nccl.lib && g_nccl.CommDestroy) g_nccl.CommDestroy((nccl_comm_t)c->nccl_comm);
  c->nccl_comm = nullptr;
}

// ---- partition geometry --------------------------------------------------------------------------------------------
// lattice planes (along the partition axis) of child `ch` that this rank owns: [lo, hi)
static void owned_planes(const Ctx* c, const Child& ch, int* lo, int* hi)
{
  const MeshDesc& m = c->mesh;
  const int pa = c->part_axis;
  const bool first = m.off[pa] == 0, last = m.off[pa] + m.n[pa] == m.gn[pa];
  *lo = first ? 0 : ch.k;                                   // planes 0..k-1 belong to the ghost layer's owner
  *hi = last ? ch.k * m.n[pa] + 1 : ch.k * m.n[pa];         // the top plane belongs to the upper neighbour
}

__global__ void k_owned_flags(int64_t size, int64_t offset, const int32_t* __restrict__ dof2lat, int64_t plane_pts, int lo, int hi, uint8_t* owned)
{
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= size) return;
  const int pl = (int)(dof2lat[g] / plane_pts);
  owned[offset + g] = (pl >= lo && pl < hi) ? 1 : 0;
}

__global__ void k_count_u8(const uint8_t* __restrict__ f, int64_t n, unsigned long long* out)
{
  unsigned long long a = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a += f[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0 && a) atomicAdd(out, a);
}

static int64_t plane_points(const Ctx* c, const Child& ch)
{
  int64_t p = 1;
  for (int d = 0; d < c->part_axis; ++d) p *= ch.lat_n[d];
  return p;
}

// called from mdb_finalize when the mesh is partitioned
void setup_partition(Ctx* c)
{
  const MeshDesc& m = c->mesh;
  int pa = -1;
  for (int d = 0; d < m.dim; ++d)
    if (m.n[d] != m.gn[d]) { MDB_REQUIRE(pa < 0, MDB_EINVAL, "partition: only slabs (one cut axis) are supported"); pa = d; }
  if (pa < 0) pa = m.dim - 1;                               // a one-rank "partition"
  MDB_REQUIRE(pa == m.dim - 1, MDB_EINVAL, "partition: the cut axis must be the slowest axis (dim-1)");
  c->part_axis = pa;
  dfree(c->d_owned);
  c->d_owned = dalloc<uint8_t>(c->ndof);
  c->child_owned.assign(c->children.size(), 0);
  unsigned long long* d_cnt = dalloc<unsigned long long>(1);
  c->owned_rows = 0;
  c->halo_cap = 0;
  for (size_t i = 0; i < c->children.size(); ++i) {
    const Child& ch = c->children[i];
    c->halo_cap += plane_points(c, ch) * ch.k;
    if (ch.size == 0) continue;
    int lo, hi; owned_planes(c, ch, &lo, &hi);
    MDB_LAUNCH(c, k_owned_flags, cdiv(ch.size, 256), 256, 0, ch.size, ch.offset, ch.d_dof2lat, plane_points(c, ch), lo, hi, c->d_owned);
    MDB_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), c->stream));
    MDB_LAUNCH(c, k_count_u8, 148 * 4, 256, 0, c->d_owned + ch.offset, ch.size, d_cnt);
    unsigned long long h = 0;
    MDB_CUDA(cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    c->child_owned[i] = (int64_t)h;
    c->owned_rows += (int64_t)h;
  }
  dfree(d_cnt);
  // receive window: [parity 2][side 2][halo_cap] doubles, then 2 sequence numbers (one per side), 64-byte aligned
  if (c->d_halo_recv) { cudaFree(c->d_halo_recv); c->d_halo_recv = nullptr; }
  c->halo_recv_n = 4 * c->halo_cap + 8;
  c->d_halo_recv = dalloc<double>(c->halo_recv_n);
  MDB_CUDA(cudaMemsetAsync(c->d_halo_recv, 0, c->halo_recv_n * sizeof(double), c->stream));
  if (!c->d_halo_counter) c->d_halo_counter = dalloc<unsigned int>(4);
  MDB_CUDA(cudaMemsetAsync(c->d_halo_counter, 0, 4 * sizeof(unsigned int), c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->halo_seq = 0;
}

// ---- halo exchange ----------------------------------------------------------------------------------------------------
struct HaloChild { int64_t offset, plane_pts, slot_base; const int32_t* lat2dof; int k; int send_lo_plane, send_hi_plane0, ghost_hi_plane; };
struct HaloSpec {
  int nchild; HaloChild ch[MDB_MAX_CHILDREN];
  int64_t cap;
  double* peer[2];          // windows of the lower / upper neighbour (null: no such neighbour)
  double* mine;             // own window
  unsigned long long seq;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Push: direction 0 sends my first owned plane to the lower neighbour (it lands in that rank's side-1 = "from above" slot);
// direction 1 sends my last k owned planes to the upper neighbour (side-0 = "from below" slot).  The last block to finish
// publishes the sequence number in both neighbours' windows.
__global__ void __launch_bounds__(256) k_halo_push(HaloSpec H, const double* __restrict__ x, unsigned int* counter)
{
  const int parity = (int)(H.seq & 1);
  for (int dir = 0; dir < 2; ++dir) {
    double* win = H.peer[dir];
    if (!win) continue;
    double* dst = win + ((int64_t)parity * 2 + (1 - dir)) * H.cap;
    for (int ci = 0; ci < H.nchild; ++ci) {
      const HaloChild& ch = H.ch[ci];
      const int np = dir == 0 ? 1 : ch.k;
      const int p0 = dir == 0 ? ch.send_lo_plane : ch.send_hi_plane0;
      const int64_t total = ch.plane_pts * np;
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t g = ch.lat2dof[(int64_t)p0 * ch.plane_pts + i];
        dst[ch.slot_base + i] = g >= 0 ? x[ch.offset + g] : 0.0;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int dir = 0; dir < 2; ++dir)
        if (H.peer[dir]) st_release_sys(reinterpret_cast<unsigned long long*>(H.peer[dir] + 4 * H.cap) + (1 - dir), H.seq);
    }
  }
}

// Wait + unpack: acquire the neighbours' sequence numbers, then copy the window into the ghost entries of x.  The wait is
// bounded (about two seconds of SM clock) so that a lost peer can never hang the device; a timeout raises *err.
__global__ void __launch_bounds__(256) k_halo_unpack(HaloSpec H, double* __restrict__ x, int has_lower, int has_upper, int* err, long long timeout_clocks,
                                                     double* done)
{
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(H.mine + 4 * H.cap);
    const long long t0 = clock64();
    bool good = true;
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? has_lower : has_upper)) continue;
      while (ld_acquire_sys(flags + side) < H.seq) {
        if (clock64() - t0 > timeout_clocks) { good = false; break; }
        __nanosleep(100);
      }
    }
    ok = good ? 1 : 0;
    if (!good) { atomicExch(err, 1); if (done) *done = 1.0; }        // a Krylov loop on this stream stops instead of iterating on stale ghosts
  }
  __syncthreads();
  if (!ok) return;
  const int parity = (int)(H.seq & 1);
  for (int side = 0; side < 2; ++side) {
    if (!(side == 0 ? has_lower : has_upper)) continue;
    const double* src = H.mine + ((int64_t)parity * 2 + side) * H.cap;
    for (int ci = 0; ci < H.nchild; ++ci) {
      const HaloChild& ch = H.ch[ci];
      const int np = side == 0 ? ch.k : 1;
      const int p0 = side == 0 ? 0 : ch.ghost_hi_plane;
      const int64_t total = ch.plane_pts * np;
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t g = ch.lat2dof[(int64_t)p0 * ch.plane_pts + i];
        if (g >= 0) x[ch.offset + g] = src[ch.slot_base + i];
      }
    }
  }
}


// ---- halo exchange over index lists (unstructured partitions) ------------------------------------------------------------------------
// A recursive-bisection partition of a gmsh mesh has no planes: every rank owns an arbitrary subset of its local DOFs and
// has up to MDB_MAX_NEIGH neighbours.  mdb_partition_lists declares, per neighbour, which owned entries are sent and which ghost entries
// are received (same order on both sides); the window of a rank is the concatenation of its receive blocks, double-buffered by the
// parity of the exchange number, followed by one sequence number per neighbour.  Same protocol as the plane exchange above: peer
// stores, one release-store of the sequence number per neighbour by the last block, acquire-spin + unpack on the receiver.
#define MDB_MAX_NEIGH 16
struct HaloListsHost {
  int nn = 0;
  int ranks[MDB_MAX_NEIGH];
  int64_t send_ptr[MDB_MAX_NEIGH + 1], recv_ptr[MDB_MAX_NEIGH + 1];
  int32_t* d_send_idx = nullptr; int32_t* d_recv_idx = nullptr;
  double* peer[MDB_MAX_NEIGH]; int64_t peer_cap[MDB_MAX_NEIGH], peer_off[MDB_MAX_NEIGH]; int peer_slot[MDB_MAX_NEIGH];
  bool imported = false;
};
struct HaloL {
  int nn; int64_t send_ptr[MDB_MAX_NEIGH + 1], recv_ptr[MDB_MAX_NEIGH + 1];
  double* peer[MDB_MAX_NEIGH]; int64_t peer_cap[MDB_MAX_NEIGH], peer_off[MDB_MAX_NEIGH]; int peer_slot[MDB_MAX_NEIGH];
  const int32_t* send_idx; const int32_t* recv_idx; double* mine; int64_t cap; unsigned long long seq;
};
__global__ void __launch_bounds__(256) k_halo_push_lists(HaloL H, const double* __restrict__ x, unsigned int* counter)
{
  const int parity = (int)(H.seq & 1);
  const int64_t total = H.send_ptr[H.nn];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int nb = 0;
    while (i >= H.send_ptr[nb + 1]) ++nb;
    H.peer[nb][(int64_t)parity * H.peer_cap[nb] + H.peer_off[nb] + (i - H.send_ptr[nb])] = x[H.send_idx[i]];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {
      *counter = 0;
      __threadfence_system();
      for (int nb = 0; nb < H.nn; ++nb) st_release_sys(reinterpret_cast<unsigned long long*>(H.peer[nb] + 2 * H.peer_cap[nb]) + H.peer_slot[nb], H.seq);
    }
  }
}
__global__ void __launch_bounds__(256) k_halo_unpack_lists(HaloL H, double* __restrict__ x, int* err, long long timeout_clocks, double* done)
{
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(H.mine + 2 * H.cap);
    const long long t0 = clock64();
    bool good = true;
    for (int nb = 0; nb < H.nn && good; ++nb)
      while (ld_acquire_sys(flags + nb) < H.seq) {
        if (clock64() - t0 > timeout_clocks) { good = false; break; }
        __nanosleep(100);
      }
    ok = good ? 1 : 0;
    if (!good) { atomicExch(err, 1); if (done) *done = 1.0; }
  }
  __syncthreads();
  if (!ok) return;
  const double* src = H.mine + (int64_t)(H.seq & 1) * H.cap;
  const int64_t total = H.recv_ptr[H.nn];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) x[H.recv_idx[i]] = src[i];
}
static HaloL halo_lists_spec(Ctx* c)
{
  const HaloListsHost* L = (const HaloListsHost*)c->halo_lists;
  HaloL H; H.nn = L->nn;
  for (int i = 0; i <= L->nn; ++i) { H.send_ptr[i] = L->send_ptr[i]; H.recv_ptr[i] = L->recv_ptr[i]; }
  for (int i = 0; i < L->nn; ++i) { H.peer[i] = L->peer[i]; H.peer_cap[i] = L->peer_cap[i]; H.peer_off[i] = L->peer_off[i]; H.peer_slot[i] = L->peer_slot[i]; }
  H.send_idx = L->d_send_idx; H.recv_idx = L->d_recv_idx; H.mine = c->d_halo_recv; H.cap = c->halo_cap; H.seq = c->halo_seq;
  return H;
}
void halo_lists_free(Ctx* c)
{
  HaloListsHost* L = (HaloListsHost*)c->halo_lists;
  if (!L) return;
  if (L->d_send_idx) cudaFree(L->d_send_idx);
  if (L->d_recv_idx) cudaFree(L->d_recv_idx);
  delete L; c->halo_lists = nullptr;
}

static long long halo_timeout_clocks(const Ctx* c);

static HaloSpec halo_spec(Ctx* c)
{
  const MeshDesc& m = c->mesh;
  const int pa = c->part_axis;
  HaloSpec H; H.nchild = (int)c->children.size(); H.cap = c->halo_cap; H.mine = c->d_halo_recv;
  H.peer[0] = c->peer_lower; H.peer[1] = c->peer_upper; H.seq = c->halo_seq;
  int64_t base = 0;
  for (int i = 0; i < H.nchild; ++i) {
    const Child& ch = c->children[i];
    HaloChild& h = H.ch[i];
    h.offset = ch.offset; h.plane_pts = plane_points(c, ch); h.slot_base = base; h.lat2dof = ch.d_lat2dof; h.k = ch.k;
    h.send_lo_plane = ch.k;                               // first owned plane of a rank that has a lower neighbour
    h.send_hi_plane0 = ch.k * m.n[pa] - ch.k;             // last k owned planes of a rank that has an upper neighbour
    h.ghost_hi_plane = ch.k * m.n[pa];
    base += h.plane_pts * ch.k;
  }
  return H;
}

void halo_push(Ctx* c, const double* d_x)
{
  MDB_REQUIRE(c->d_halo_recv, MDB_ESTATE, "halo exchange on an unpartitioned context");
  c->halo_seq++;
  if (c->halo_lists) {
    const HaloListsHost* L = (const HaloListsHost*)c->halo_lists;
    if (L->nn == 0) return;
    MDB_REQUIRE(L->imported, MDB_ESTATE, "halo exchange before mdb_halo_import_lists");
    const HaloL H = halo_lists_spec(c);
    int grid = cdiv(H.send_ptr[H.nn], 256); if (grid > 148) grid = 148; if (grid < 1) grid = 1;
    MDB_LAUNCH(c, k_halo_push_lists, grid, 256, 0, H, d_x, c->d_halo_counter);
    return;
  }
  if (!c->peer_lower && !c->peer_upper) return;
  HaloSpec H = halo_spec(c);
  int64_t work = 0;
  for (int i = 0; i < H.nchild; ++i) work += H.ch[i].plane_pts * H.ch[i].k;
  int grid = cdiv(work, 256); if (grid > 148) grid = 148; if (grid < 1) grid = 1;
  MDB_LAUNCH(c, k_halo_push, grid, 256, 0, H, d_x, c->d_halo_counter);
}

static long long halo_timeout_clocks(const Ctx* c)
{
  double s = c->halo_timeout_s;
  if (s <= 0.0) { const char* e = getenv("MDB_HALO_TIMEOUT_S"); s = e ? atof(e) : 30.0; }
  if (s < 1e-3) s = 1e-3;
  return (long long)(s * 1.9e9);                         // SM clocks at ~1.9 GHz
}

void halo_wait(Ctx* c, double* d_x, double* d_done)
{
  MDB_REQUIRE(c->d_halo_recv, MDB_ESTATE, "halo exchange on an unpartitioned context");
  if (c->halo_lists) {
    const HaloListsHost* L = (const HaloListsHost*)c->halo_lists;
    if (L->nn == 0) return;
    const HaloL Hl = halo_lists_spec(c);
    int gridl = cdiv(Hl.recv_ptr[Hl.nn], 256); if (gridl > 64) gridl = 64; if (gridl < 1) gridl = 1;
    MDB_LAUNCH(c, k_halo_unpack_lists, gridl, 256, 0, Hl, d_x, (int*)(c->d_halo_counter + 2), halo_timeout_clocks(c), d_done);
    return;
  }
  if (!c->peer_lower && !c->peer_upper) return;
  HaloSpec H = halo_spec(c);
  int64_t work = 0;
  for (int i = 0; i < H.nchild; ++i) work += H.ch[i].plane_pts * H.ch[i].k;
  int grid = cdiv(work, 256); if (grid > 64) grid = 64; if (grid < 1) grid = 1;     // every block spins on the flags: keep the grid co-resident
  MDB_LAUNCH(c, k_halo_unpack, grid, 256, 0, H, d_x, c->peer_lower ? 1 : 0, c->peer_upper ? 1 : 0, (int*)(c->d_halo_counter + 2),
             halo_timeout_clocks(c), d_done);
}

void halo_exchange(Ctx* c, double* d_x, double* d_done)
{
  halo_push(c, d_x);
  halo_wait(c, d_x, d_done);
}

bool halo_timed_out(Ctx* c)
{
  if (c->sring && ((ScalarRing*)c->sring)->d_err) {
    ScalarRing* R = (ScalarRing*)c->sring;
    int e = 0;
    MDB_CUDA(cudaMemcpyAsync(&e, R->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
    if (e != 0) { MDB_CUDA(cudaMemsetAsync(R->d_err, 0, sizeof(int), c->stream)); MDB_CUDA(cudaStreamSynchronize(c->stream)); return true; }
  }
  if (!c->d_halo_counter) return false;
  int e = 0;
  MDB_CUDA(cudaMemcpyAsync(&e, c->d_halo_counter + 2, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  // report once: a transient stall of a neighbour (a profiler replay, a descheduled rank) must not fail every later exchange
  if (e != 0) { MDB_CUDA(cudaMemsetAsync(c->d_halo_counter + 2, 0, sizeof(int), c->stream)); MDB_CUDA(cudaStreamSynchronize(c->stream)); }
  return e != 0;
}

} // namespace mdb

using namespace mdb;
struct mdb_ctx : mdb::Ctx {};

#define MG_BEGIN(ctx) if (!(ctx)) return MDB_EINVAL; try { cudaSetDevice((ctx)->device);
#define MG_END(ctx) return MDB_OK; } catch (mdb::Error& e) { (ctx)->err = e.what(); return e.code; } catch (std::exception& e) { (ctx)->err = e.what(); return MDB_EINVAL; }

struct HaloBlob { cudaIpcMemHandle_t ipc; uint64_t raw; uint32_t token; int32_t device; };
// "same process" is decided by a random per-process token, not by the pid alone (pids repeat across pid namespaces)
static uint32_t process_token()
{
  static uint32_t t = 0;
  if (!t) {
    FILE* f = fopen("/dev/urandom", "rb");
    if (f) { if (fread(&t, sizeof(t), 1, f) != 1) t = 0; fclose(f); }
    if (!t) t = (uint32_t)getpid() * 2654435761u ^ (uint32_t)((uintptr_t)&t & 0xFFFFFFFFu) ^ 0x9E3779B9u;
  }
  return t;
}
static_assert(sizeof(HaloBlob) <= 80, "halo blob must fit the 80-byte ABI buffer");

extern "C" {

int mdb_nccl_unique_id(void* id128)
{
  try { nccl_uid_t id; MDB_NCCL(nccl().GetUniqueId(&id)); std::memcpy(id128, &id, 128); return MDB_OK; }
  catch (std::exception&) { return MDB_ENCCL; }
}

int mdb_comm_init(mdb_ctx* c, int nranks, int rank, const void* id128)
{
  MG_BEGIN(c)
  MDB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, MDB_EINVAL, "mdb_comm_init: bad rank");
  c->nranks = nranks; c->rank = rank;
  if (nranks > 1) {
    MDB_REQUIRE(id128, MDB_EINVAL, "mdb_comm_init: unique id required");
    nccl_uid_t id; std::memcpy(&id, id128, 128);
    nccl_comm_t comm = nullptr;
    if (c->nccl_comm) { nccl().CommDestroy((nccl_comm_t)c->nccl_comm); c->nccl_comm = nullptr; }      // re-initialisation
    MDB_NCCL(nccl().CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = comm;
  }
  MG_END(c)
}

int mdb_set_halo_timeout(mdb_ctx* c, double seconds)
{
  if (!c) return MDB_EINVAL;
  c->halo_timeout_s = seconds;
  return MDB_OK;
}

int mdb_halo_export(mdb_ctx* c, void* blob80, int64_t* window_bytes)
{
  MG_BEGIN(c)
  MDB_REQUIRE(c->finalized && c->d_halo_recv, MDB_ESTATE, "mdb_halo_export: finalize a partitioned mesh first");
  HaloBlob b; std::memset(&b, 0, sizeof(b));
  MDB_CUDA(cudaIpcGetMemHandle(&b.ipc, c->d_halo_recv));
  b.raw = (uint64_t)(uintptr_t)c->d_halo_recv; b.device = c->device; b.token = process_token();
  std::memset(blob80, 0, 80);
  std::memcpy(blob80, &b, sizeof(b));
  if (window_bytes) *window_bytes = c->halo_recv_n * (int64_t)sizeof(double);
  MG_END(c)
}

static double* open_peer(mdb_ctx* c, const void* blob80);
int mdb_scalars_export(mdb_ctx* c, void* blob80)
{
  MG_BEGIN(c)
  MDB_REQUIRE(c->nranks > 1, MDB_ESTATE, "mdb_scalars_export: call mdb_comm_init with more than one rank first");
  MDB_REQUIRE(c->nranks <= SR_MAXR, MDB_EINVAL, "mdb_scalars_export: at most 16 ranks");
  sring_release(c);
  ScalarRing* R = new ScalarRing(); c->sring = R;
  R->nranks = c->nranks; R->rank = c->rank;
  void* p = nullptr; MDB_CUDA(cudaMalloc(&p, SR_WINDOW_BYTES)); R->mine = (double*)p;
  MDB_CUDA(cudaMemset(R->mine, 0, SR_WINDOW_BYTES));
  R->d_err = dalloc<int>(1); MDB_CUDA(cudaMemset(R->d_err, 0, sizeof(int)));
  HaloBlob b; std::memset(&b, 0, sizeof(b));
  MDB_CUDA(cudaIpcGetMemHandle(&b.ipc, R->mine));
  b.raw = (uint64_t)(uintptr_t)R->mine; b.device = c->device; b.token = process_token();
  std::memset(blob80, 0, 80);
  std::memcpy(blob80, &b, sizeof(b));
  MG_END(c)
}

int mdb_scalars_import(mdb_ctx* c, int nranks, const void* blobs80)
{
  MG_BEGIN(c)
  ScalarRing* R = (ScalarRing*)c->sring;
  if (nranks == 0) { if (R && !c->sring_borrowed) R->connected = false; return MDB_OK; }      // disconnect: back to ncclAllReduce
  MDB_REQUIRE(R && R->mine && !c->sring_borrowed, MDB_ESTATE, "mdb_scalars_import: call mdb_scalars_export first");
  MDB_REQUIRE(nranks == c->nranks && blobs80, MDB_EINVAL, "mdb_scalars_import: one descriptor per rank of the communicator");
  for (int r = 0; r < nranks; ++r) R->win[r] = r == c->rank ? R->mine : open_peer(c, (const char*)blobs80 + 80 * (size_t)r);
  R->connected = true;
  MG_END(c)
}

int64_t mdb_scalars_exchanges(mdb_ctx* c)
{
  if (!c || !sring_connected(c)) return 0;
  return (int64_t)((ScalarRing*)c->sring)->seq;
}

static double* open_peer(mdb_ctx* c, const void* blob80)
{
  HaloBlob b; std::memcpy(&b, blob80, sizeof(b));
  if (b.token == process_token()) {                        // a context of this very process: plain pointer (+ peer access if another device)
    if (b.device != c->device) { cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0); if (e != cudaSuccess) cudaGetLastError(); }
    return (double*)(uintptr_t)b.raw;
  }
  void* p = nullptr;
  MDB_CUDA(cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
  c->ipc_opened.push_back(p);
  return (double*)p;
}

int mdb_halo_import(mdb_ctx* c, int lower_rank, const void* lower_blob80, int upper_rank, const void* upper_blob80)
{
  MG_BEGIN(c)
  MDB_REQUIRE(c->finalized && c->d_halo_recv, MDB_ESTATE, "mdb_halo_import: finalize a partitioned mesh first");
  const MeshDesc& m = c->mesh; const int pa = c->part_axis;
  const bool first = m.off[pa] == 0, last = m.off[pa] + m.n[pa] == m.gn[pa];
  MDB_REQUIRE(first == (lower_blob80 == nullptr), MDB_EINVAL, "mdb_halo_import: lower neighbour does not match the partition");
  MDB_REQUIRE(last == (upper_blob80 == nullptr), MDB_EINVAL, "mdb_halo_import: upper neighbour does not match the partition");
  c->lower_rank = lower_blob80 ? lower_rank : -1; c->upper_rank = upper_blob80 ? upper_rank : -1;
  c->peer_lower = lower_blob80 ? open_peer(c, lower_blob80) : nullptr;
  c->peer_upper = upper_blob80 ? open_peer(c, upper_blob80) : nullptr;
  MG_END(c)
}


int mdb_partition_lists(mdb_ctx* c, const uint8_t* owned_flags, int nneigh, const int* neigh_ranks, const int64_t* send_ptr, const int32_t* send_idx,
                        const int64_t* recv_ptr, const int32_t* recv_idx)
{
  MG_BEGIN(c)
  MDB_REQUIRE(c->finalized && (!c->mesh.partitioned || c->part_axis < 0), MDB_ESTATE, "mdb_partition_lists: after mdb_finalize, on a mesh without a plane (slab) partition");
  MDB_REQUIRE(owned_flags && nneigh >= 0 && nneigh <= MDB_MAX_NEIGH, MDB_EINVAL, "mdb_partition_lists: 0..16 neighbours");
  MDB_REQUIRE(nneigh == 0 || (neigh_ranks && send_ptr && recv_ptr), MDB_EINVAL, "mdb_partition_lists: null list");
  halo_lists_free(c);
  HaloListsHost* L = new HaloListsHost(); c->halo_lists = L;
  L->nn = nneigh; L->send_ptr[0] = L->recv_ptr[0] = 0;
  for (int i = 0; i < nneigh; ++i) {
    L->ranks[i] = neigh_ranks[i]; L->send_ptr[i + 1] = send_ptr[i + 1]; L->recv_ptr[i + 1] = recv_ptr[i + 1];
    MDB_REQUIRE(send_ptr[i + 1] >= send_ptr[i] && recv_ptr[i + 1] >= recv_ptr[i] && send_ptr[0] == 0 && recv_ptr[0] == 0, MDB_EINVAL, "mdb_partition_lists: bad list pointers");
    L->peer[i] = nullptr; L->peer_cap[i] = 0; L->peer_off[i] = 0; L->peer_slot[i] = 0;
  }
  const int64_t ns = L->send_ptr[nneigh], nr = L->recv_ptr[nneigh];
  for (int64_t i = 0; i < ns; ++i) MDB_REQUIRE(send_idx[i] >= 0 && send_idx[i] < c->ndof && owned_flags[send_idx[i]], MDB_EINVAL, "mdb_partition_lists: a sent entry is not an owned DOF");
  for (int64_t i = 0; i < nr; ++i) MDB_REQUIRE(recv_idx[i] >= 0 && recv_idx[i] < c->ndof && !owned_flags[recv_idx[i]], MDB_EINVAL, "mdb_partition_lists: a received entry is not a ghost DOF");
  L->d_send_idx = dalloc<int32_t>(ns); L->d_recv_idx = dalloc<int32_t>(nr);
  if (ns) MDB_CUDA(cudaMemcpyAsync(L->d_send_idx, send_idx, ns * 4, cudaMemcpyHostToDevice, c->stream));
  if (nr) MDB_CUDA(cudaMemcpyAsync(L->d_recv_idx, recv_idx, nr * 4, cudaMemcpyHostToDevice, c->stream));
  dfree(c->d_owned);
  c->d_owned = dalloc<uint8_t>(c->ndof);
  MDB_CUDA(cudaMemcpyAsync(c->d_owned, owned_flags, c->ndof, cudaMemcpyHostToDevice, c->stream));
  c->child_owned.assign(c->children.size(), 0); c->owned_rows = 0;
  for (size_t k = 0; k < c->children.size(); ++k) {
    int64_t n = 0;
    for (int64_t g = c->children[k].offset; g < c->children[k].offset + c->children[k].size; ++g) n += owned_flags[g] ? 1 : 0;
    c->child_owned[k] = n; c->owned_rows += n;
  }
  // receive window: [parity 2][cap] doubles, then MDB_MAX_NEIGH sequence numbers
  c->halo_cap = nr > 0 ? nr : 1;
  if (c->d_halo_recv) { cudaFree(c->d_halo_recv); c->d_halo_recv = nullptr; }
  c->halo_recv_n = 2 * c->halo_cap + MDB_MAX_NEIGH;
  c->d_halo_recv = dalloc<double>(c->halo_recv_n);
  MDB_CUDA(cudaMemsetAsync(c->d_halo_recv, 0, c->halo_recv_n * sizeof(double), c->stream));
  if (!c->d_halo_counter) c->d_halo_counter = dalloc<unsigned int>(4);
  MDB_CUDA(cudaMemsetAsync(c->d_halo_counter, 0, 4 * sizeof(unsigned int), c->stream));
  MDB_CUDA(cudaStreamSynchronize(c->stream));
  c->halo_seq = 0;
  for (Matrix& M : c->mats) { M.dinv_valid = false; M.ilu_valid = false; }
  MG_END(c)
}

int mdb_halo_import_lists(mdb_ctx* c, const void* blobs80, const int64_t* peer_cap, const int64_t* peer_offset, const int* peer_slot)
{
  MG_BEGIN(c)
  HaloListsHost* L = (HaloListsHost*)c->halo_lists;
  MDB_REQUIRE(L, MDB_ESTATE, "mdb_halo_import_lists: call mdb_partition_lists first");
  for (int i = 0; i < L->nn; ++i) {
    MDB_REQUIRE(peer_cap[i] >= 1 && peer_offset[i] >= 0 && peer_offset[i] + (L->send_ptr[i + 1] - L->send_ptr[i]) <= peer_cap[i] && peer_slot[i] >= 0 && peer_slot[i] < MDB_MAX_NEIGH,
                MDB_EINVAL, "mdb_halo_import_lists: my block does not fit the neighbour's window");
    L->peer[i] = open_peer(c, (const char*)blobs80 + 80 * (size_t)i);
    L->peer_cap[i] = peer_cap[i]; L->peer_off[i] = peer_offset[i]; L->peer_slot[i] = peer_slot[i];
  }
  L->imported = true;
  MG_END(c)
}

int mdb_halo_push(mdb_ctx* c, const double* x)
{
  MG_BEGIN(c)
  MDB_REQUIRE(is_device_ptr(x), MDB_EINVAL, "mdb_halo_push: x must be a device vector");
  halo_push(c, x);
  MG_END(c)
}

int mdb_halo_wait(mdb_ctx* c, double* x)
{
  MG_BEGIN(c)
  MDB_REQUIRE(is_device_ptr(x), MDB_EINVAL, "mdb_halo_wait: x must be a device vector");
  halo_wait(c, x, nullptr);
  MDB_REQUIRE(!halo_timed_out(c), MDB_ENCCL, "halo exchange timed out waiting for a neighbour");
  MG_END(c)
}

int mdb_owned_rows(mdb_ctx* c, int64_t* owned, int64_t* local_total)
{
  if (!c) return MDB_EINVAL;
  if (owned) *owned = c->d_owned ? c->owned_rows : c->ndof;
  if (local_total) *local_total = c->ndof;
  return MDB_OK;
}

int mdb_owned_counts(mdb_ctx* c, int64_t* per_child)
{
  if (!c || !c->finalized) return MDB_EINVAL;
  for (size_t i = 0; i < c->children.size(); ++i) per_child[i] = c->d_owned ? c->child_owned[i] : c->children[i].size;
  return MDB_OK;
}

int mdb_get_ownership(mdb_ctx* c, const int64_t* child_global_base, uint8_t* owned_flags, int64_t* global_index)
{
  MG_BEGIN(c)
  MDB_REQUIRE(c->finalized, MDB_ESTATE, "mdb_get_ownership: not finalized");
  std::vector<uint8_t> f(c->ndof, 1);
  if (c->d_owned) {
    MDB_CUDA(cudaMemcpyAsync(f.data(), c->d_owned, c->ndof, cudaMemcpyDeviceToHost, c->stream));
    MDB_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (owned_flags) std::memcpy(owned_flags, f.data(), c->ndof);
  if (global_index) {
    // owned dofs of a child keep their relative order in the global (lexicographic, cut axis slowest) numbering.  True for Q1 blocks
    // only: a Qk (k > 1) block is numbered entity group by entity group, so its global numbers are not contiguous per rank
    if (child_global_base)
      for (const Child& ch : c->children)
        MDB_REQUIRE(ch.k == 1, MDB_EINVAL, "mdb_get_ownership: global indices by rank offsets exist for Q1 blocks only (Qk: map through mdb_get_dofmap)");
    for (size_t i = 0; i < c->children.size(); ++i) {
      const Child& ch = c->children[i];
      int64_t next = child_global_base ? child_global_base[i] : ch.offset;
      for (int64_t g = ch.offset; g < ch.offset + ch.size; ++g) global_index[g] = f[g] ? next++ : -1;
    }
  }
  MG_END(c)
}

} // extern "C"
