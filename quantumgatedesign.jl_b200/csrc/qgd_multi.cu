// qgd_multi.cu -- NCCL plumbing of the multi-GPU entry points (include/qgd_b200.h, "multi-GPU inside the library").
//
// The reference parallelises over the independent initial-condition columns with Threads.@threads
// (src/forward_evolution.jl:48,332); the columns couple only through dot(final_state, R), dot(final_state, T) in
// compute_terminal_condition (src/eval_grad_discrete_adjoint.jl:27-28) and through the serial gradient sum
// (:150-157).  Here the columns are sharded over GPUs and those two couplings are NCCL collectives on device buffers,
// enqueued on the sweep stream between the kernels -- no host staging.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy the hosting process already loaded, e.g. the one bundled
// with PyTorch under torchrun, else the system library), so that a single-GPU user of libqgd_b200.so needs no NCCL.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "qgd_host.h"

namespace qgd_nccl {

namespace {
// Minimal declarations of the stable NCCL 2.x C API (nccl.h: ncclUniqueId is 128 opaque bytes, ncclFloat64 = 8,
// ncclSum = 0, ncclSuccess = 0).
struct UniqueId { char internal[128]; };
typedef void* Comm;
typedef int (*GetUniqueId_t)(UniqueId*);
typedef int (*CommInitRank_t)(Comm*, int, UniqueId, int);
typedef int (*CommInitAll_t)(Comm*, int, const int*);
typedef int (*CommDestroy_t)(Comm);
typedef int (*AllReduce_t)(const void*, void*, size_t, int, int, Comm, cudaStream_t);
typedef int (*Group_t)(void);
typedef const char* (*GetErrorString_t)(int);
typedef int (*GetVersion_t)(int*);

struct Api {
  void* lib = nullptr;
  GetUniqueId_t GetUniqueId = nullptr;
  CommInitRank_t CommInitRank = nullptr;
  CommInitAll_t CommInitAll = nullptr;
  CommDestroy_t CommDestroy = nullptr;
  AllReduce_t AllReduce = nullptr;
  Group_t GroupStart = nullptr, GroupEnd = nullptr;
  GetErrorString_t GetErrorString = nullptr;
  GetVersion_t GetVersion = nullptr;
  std::string path, error;
};
Api g_api;
std::mutex g_mu;
std::string g_override;

bool load_locked() {
  if (g_api.lib) return true;
  const char* names[] = {g_override.empty() ? nullptr : g_override.c_str(), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    void* l = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (l) { g_api.lib = l; g_api.path = nm; break; }
    g_api.error = dlerror();
  }
  if (!g_api.lib) return false;
  auto sym = [&](const char* s) { void* p = dlsym(g_api.lib, s); if (!p) g_api.error = std::string("missing NCCL symbol ") + s; return p; };
  g_api.GetUniqueId = (GetUniqueId_t)sym("ncclGetUniqueId");
  g_api.CommInitRank = (CommInitRank_t)sym("ncclCommInitRank");
  g_api.CommInitAll = (CommInitAll_t)sym("ncclCommInitAll");
  g_api.CommDestroy = (CommDestroy_t)sym("ncclCommDestroy");
  g_api.AllReduce = (AllReduce_t)sym("ncclAllReduce");
  g_api.GroupStart = (Group_t)sym("ncclGroupStart");
  g_api.GroupEnd = (Group_t)sym("ncclGroupEnd");
  g_api.GetErrorString = (GetErrorString_t)sym("ncclGetErrorString");
  g_api.GetVersion = (GetVersion_t)sym("ncclGetVersion");
  if (!(g_api.GetUniqueId && g_api.CommInitRank && g_api.CommInitAll && g_api.CommDestroy && g_api.AllReduce && g_api.GroupStart &&
        g_api.GroupEnd && g_api.GetErrorString)) {
    dlclose(g_api.lib); g_api.lib = nullptr;
    return false;
  }
  return true;
}

void need() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!load_locked())
    throw QgdError(QGD_EUNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2 failed: " + g_api.error +
                                         "); multi-GPU entry points need it, single-GPU ones do not");
}
void check(int rc, const char* what) {
  if (rc != 0) throw QgdError(QGD_ECUDA, std::string(what) + ": " + g_api.GetErrorString(rc));
}
}  // namespace

void set_library(const char* path) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_override = path ? path : "";
}
int version() {
  need();
  int v = 0;
  if (g_api.GetVersion) g_api.GetVersion(&v);
  return v;
}
void unique_id(unsigned char out[128]) {
  need();
  UniqueId id;
  check(g_api.GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out, id.internal, 128);
}
void* init_rank(int nranks, int rank, const unsigned char idbytes[128]) {
  need();
  UniqueId id;
  std::memcpy(id.internal, idbytes, 128);
  Comm c = nullptr;
  check(g_api.CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
  return c;
}
void init_all(void** comms, int n, const int* devices) {
  need();
  check(g_api.CommInitAll(reinterpret_cast<Comm*>(comms), n, devices), "ncclCommInitAll");
}
void destroy(void* comm) {
  if (comm && g_api.lib) g_api.CommDestroy(comm);
}
// in-place sum of n doubles over the communicator, enqueued on `stream`
void allreduce_sum(void* comm, double* buf, size_t n, cudaStream_t stream) {
  check(g_api.AllReduce(buf, buf, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, stream), "ncclAllReduce");
}
void group_start() { check(g_api.GroupStart(), "ncclGroupStart"); }
void group_end() { check(g_api.GroupEnd(), "ncclGroupEnd"); }

}  // namespace qgd_nccl
