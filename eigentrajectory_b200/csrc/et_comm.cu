// Collective entry points of the C ABI: a thin layer over NCCL so that a host that is not
// Python can run the row-sharded stages -- local et_gram -> et_allreduce_f64 -> et_eig_jacobi_pair, or one
// et_kmeans_assign_shard -> et_allreduce_f64 -> et_kmeans_finalize per Lloyd iteration.  (The Python mirror drives the
// same collectives through torch.distributed; the fused k-means / seeding kernels need no collective call at all.)
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time dependency on it, and inside a
// process that already carries NCCL (e.g. PyTorch's bundled copy, same soname) that very copy is used.
#include <dlfcn.h>
#include <nccl.h>      // types and enums only; no symbol of libnccl is linked

#include <mutex>

#include "et_common.cuh"

namespace et {
namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) get_unique_id = nullptr;
  decltype(&ncclCommInitRank) comm_init_rank = nullptr;
  decltype(&ncclAllReduce) all_reduce = nullptr;
  decltype(&ncclCommDestroy) comm_destroy = nullptr;
  decltype(&ncclGetErrorString) error_string = nullptr;
  decltype(&ncclCommCount) comm_count = nullptr;
  decltype(&ncclCommUserRank) comm_rank = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;

const NcclApi* nccl() {
  std::call_once(g_nccl_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    NcclApi a;
    a.handle = h;
    a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(h, "ncclGetUniqueId"));
    a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(h, "ncclCommInitRank"));
    a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(dlsym(h, "ncclAllReduce"));
    a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(h, "ncclCommDestroy"));
    a.error_string = reinterpret_cast<decltype(a.error_string)>(dlsym(h, "ncclGetErrorString"));
    a.comm_count = reinterpret_cast<decltype(a.comm_count)>(dlsym(h, "ncclCommCount"));
    a.comm_rank = reinterpret_cast<decltype(a.comm_rank)>(dlsym(h, "ncclCommUserRank"));
    if (a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy && a.error_string && a.comm_count && a.comm_rank)
      g_nccl = a;
  });
  return g_nccl.handle ? &g_nccl : nullptr;
}

int nccl_fail(const NcclApi* api, ncclResult_t r, const char* what) {
  return fail(ET_ERR_NCCL, "%s: %s", what, api->error_string(r));
}

}  // namespace
}  // namespace et

using namespace et;

struct et_comm {
  ncclComm_t comm;
  int rank, nranks;
};

extern "C" {

int et_comm_unique_id(void* id_out) {
  ET_REQUIRE(id_out, ET_ERR_BADARG, "et_comm_unique_id: null pointer");
  const NcclApi* api = nccl();
  ET_REQUIRE(api, ET_ERR_NCCL, "et_comm_unique_id: libnccl.so.2 cannot be loaded");
  static_assert(sizeof(ncclUniqueId) == ET_COMM_ID_BYTES, "ET_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  ncclUniqueId id;
  ncclResult_t r = api->get_unique_id(&id);
  if (r != ncclSuccess) return nccl_fail(api, r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return ET_OK;
}

int et_comm_init(int rank, int nranks, const void* unique_id, et_comm_t* comm_out) {
  ET_REQUIRE(unique_id && comm_out, ET_ERR_BADARG, "et_comm_init: null pointer");
  ET_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, ET_ERR_BADARG, "et_comm_init: rank %d of %d", rank, nranks);
  const NcclApi* api = nccl();
  ET_REQUIRE(api, ET_ERR_NCCL, "et_comm_init: libnccl.so.2 cannot be loaded");
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r = api->comm_init_rank(&c, nranks, id, rank);      // binds the CURRENT CUDA device to this rank
  if (r != ncclSuccess) return nccl_fail(api, r, "ncclCommInitRank");
  et_comm* out = new et_comm{c, rank, nranks};
  *comm_out = out;
  return ET_OK;
}

int et_comm_rank(et_comm_t comm, int* rank, int* nranks) {
  ET_REQUIRE(comm, ET_ERR_BADARG, "et_comm_rank: null communicator");
  if (rank) *rank = comm->rank;
  if (nranks) *nranks = comm->nranks;
  return ET_OK;
}

static int all_reduce(void* buf, size_t n, ncclDataType_t type, ncclRedOp_t op, et_comm_t comm, et_stream_t stream, const char* what) {
  ET_REQUIRE(comm && (buf || n == 0), ET_ERR_BADARG, "%s: null pointer", what);
  if (n == 0) return ET_OK;
  const NcclApi* api = nccl();
  ET_REQUIRE(api, ET_ERR_NCCL, "%s: libnccl.so.2 cannot be loaded", what);
  ncclResult_t r = api->all_reduce(buf, buf, n, type, op, comm->comm, as_stream(stream));
  if (r != ncclSuccess) return nccl_fail(api, r, what);
  return ET_OK;
}

int et_allreduce_f64(double* buf, size_t n, et_comm_t comm, et_stream_t stream) {
  return all_reduce(buf, n, ncclFloat64, ncclSum, comm, stream, "et_allreduce_f64");
}

int et_allreduce_min_i64(long long* buf, size_t n, et_comm_t comm, et_stream_t stream) {
  return all_reduce(buf, n, ncclInt64, ncclMin, comm, stream, "et_allreduce_min_i64");
}

int et_comm_destroy(et_comm_t comm) {
  if (!comm) return ET_OK;
  const NcclApi* api = nccl();
  ncclResult_t r = api ? api->comm_destroy(comm->comm) : ncclSuccess;
  delete comm;
  if (api && r != ncclSuccess) return nccl_fail(api, r, "ncclCommDestroy");
  return ET_OK;
}

}  // extern "C"
