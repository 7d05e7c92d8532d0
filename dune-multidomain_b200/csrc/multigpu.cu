// multigpu.cu — multi-GPU plumbing: NCCL allreduce of Krylov scalars, halo exchange of x.
#include "common.cuh"

namespace mdb {
void allreduce_scalars(Ctx* c, double* d_vals, int n) { (void)c; (void)d_vals; (void)n; throw Error(MDB_ENCCL, "multi-GPU path not initialised"); }
void halo_exchange(Ctx* c, double* d_x) { (void)c; (void)d_x; throw Error(MDB_ENCCL, "multi-GPU path not initialised"); }
}

using namespace mdb;
struct mdb_ctx : mdb::Ctx {};
extern "C" {
int mdb_nccl_unique_id(void* id128) { (void)id128; return MDB_ENCCL; }
int mdb_comm_init(mdb_ctx* c, int nranks, int rank, const void* id128) { (void)c; (void)nranks; (void)rank; (void)id128; return MDB_ENCCL; }
int mdb_halo_export(mdb_ctx* c, void* handle64, int64_t* window_bytes) { (void)c; (void)handle64; (void)window_bytes; return MDB_ENCCL; }
int mdb_halo_import(mdb_ctx* c, int lower_rank, const void* lh, int upper_rank, const void* uh) { (void)c; (void)lower_rank; (void)lh; (void)upper_rank; (void)uh; return MDB_ENCCL; }
int mdb_owned_rows(mdb_ctx* c, int64_t* owned, int64_t* local_total) { if (!c) return MDB_EINVAL; *owned = c->ndof; *local_total = c->ndof; return MDB_OK; }
int mdb_get_ownership(mdb_ctx* c, uint8_t* owned_flags, int64_t* global_index) { (void)c; (void)owned_flags; (void)global_index; return MDB_ENCCL; }
}
