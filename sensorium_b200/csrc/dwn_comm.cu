// NCCL behind the C ABI (dwn_comm_init / dwn_allreduce_bucket): the gradient exchange of the data-parallel
// train step for hosts that do not go through torch.distributed.  NCCL is bound lazily with dlopen / dlsym - the copy the
// process has already loaded (torch's) if there is one - so the library has no link-time dependency on it and loads on
// machines without NCCL; every entry point fails loudly (-1 + dwn_last_error) when NCCL cannot be found.
#include "dwn_common.cuh"
#include <dlfcn.h>
#include <mutex>
#include <stdint.h>
#include <string.h>

namespace {
// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclBfloat16 = 9 };
enum { ncclSum = 0, ncclMax = 2, ncclAvg = 4 };
typedef int (*fn_get_unique_id)(ncclUniqueId*);
typedef int (*fn_comm_init_rank)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*fn_comm_destroy)(ncclComm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef const char* (*fn_get_error_string)(int);
typedef int (*fn_group)(void);

struct Nccl {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_get_error_string get_error_string = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
};
Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = -1, g_nranks = 0;
std::mutex g_mu;

int bind_nccl() {
  if (g_nccl.handle) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the instance torch (or the host application) already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return dwn_fail("dwn_comm: libnccl.so.2 not found (%s)", dlerror());
  Nccl n;
  n.handle = h;
  n.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
  n.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
  n.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  n.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
  n.get_error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
  n.group_start = (fn_group)dlsym(h, "ncclGroupStart");
  n.group_end = (fn_group)dlsym(h, "ncclGroupEnd");
  if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_reduce || !n.group_start || !n.group_end)
    return dwn_fail("dwn_comm: libnccl.so.2 lacks a required symbol");
  g_nccl = n;
  return 0;
}
int nccl_fail(const char* what, int rc) {
  return dwn_fail("dwn_comm: %s failed: %s", what, g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "?");
}
}  // namespace

// 128-byte rendezvous token: rank 0 creates it, the host application hands it to the other ranks (any channel)
extern "C" int dwn_comm_unique_id(void* out128) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (bind_nccl()) return -1;
  ncclUniqueId id;
  const int rc = g_nccl.get_unique_id(&id);
  if (rc != ncclSuccess) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(out128, &id, sizeof(id));
  return 0;
}

// one communicator per process (one process per GPU); the CUDA device must be current
extern "C" int dwn_comm_init(int rank, int nranks, const void* unique_id128) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (bind_nccl()) return -1;
  DWN_REQUIRE(g_comm == nullptr, "dwn_comm_init: communicator already initialised (rank %d of %d)", g_rank, g_nranks);
  DWN_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "dwn_comm_init: bad rank %d / nranks %d", rank, nranks);
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  const int rc = g_nccl.comm_init_rank(&g_comm, nranks, id, rank);
  if (rc != ncclSuccess) { g_comm = nullptr; return nccl_fail("ncclCommInitRank", rc); }
  g_rank = rank;
  g_nranks = nranks;
  return 0;
}

// in-place all-reduce of one gradient bucket on comm_stream, asynchronous; avg != 0: mean over ranks (DDP semantics),
// op_max != 0: maximum (the per-mouse has-grad flags)
extern "C" int dwn_allreduce_bucket(void* ptr, long count, int dtype, int avg, int op_max, void* comm_stream) {
  DWN_REQUIRE(g_comm != nullptr, "dwn_allreduce_bucket: dwn_comm_init has not been called");
  DWN_REQUIRE(count >= 0, "dwn_allreduce_bucket: negative count");
  int ty;
  if (dtype == DWN_DT_F32) ty = ncclFloat32;
  else if (dtype == DWN_DT_BF16) ty = ncclBfloat16;
  else if (dtype == 2) ty = ncclInt32;
  else return dwn_fail("dwn_allreduce_bucket: dtype %d unsupported (0 = fp32, 1 = bf16, 2 = int32)", dtype);
  const int op = op_max ? ncclMax : (avg ? ncclAvg : ncclSum);
  const int rc = g_nccl.all_reduce(ptr, ptr, (size_t)count, ty, op, g_comm, (cudaStream_t)comm_stream);
  if (rc != ncclSuccess) return nccl_fail("ncclAllReduce", rc);
  return 0;
}

// several buckets as one NCCL group (one launch): call between dwn_comm_group_begin / dwn_comm_group_end
extern "C" int dwn_comm_group_begin(void) {
  DWN_REQUIRE(g_comm != nullptr, "dwn_comm_group_begin: dwn_comm_init has not been called");
  const int rc = g_nccl.group_start();
  return rc == ncclSuccess ? 0 : nccl_fail("ncclGroupStart", rc);
}
extern "C" int dwn_comm_group_end(void) {
  DWN_REQUIRE(g_comm != nullptr, "dwn_comm_group_end: dwn_comm_init has not been called");
  const int rc = g_nccl.group_end();
  return rc == ncclSuccess ? 0 : nccl_fail("ncclGroupEnd", rc);
}

extern "C" int dwn_comm_destroy(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g_comm) return 0;
  const int rc = g_nccl.comm_destroy(g_comm);
  g_comm = nullptr;
  g_rank = -1;
  g_nranks = 0;
  return rc == ncclSuccess ? 0 : nccl_fail("ncclCommDestroy", rc);
}
