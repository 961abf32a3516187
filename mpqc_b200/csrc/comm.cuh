// NCCL binding and the communicator object of the (T) path (include/mpqc_t.h: mpqc_t_comm_*).
//
// The path has two exchanges over NVLink 5 / NVSwitch, both on the library's own streams:
//   * input replication: every rank copies 1/N of each host tensor over ITS PCIe link and one ncclAllGather
//     completes the tensor on all GPUs (replaces N identical 19 GB host->device uploads, which contend for one
//     host's memory bandwidth);
//   * the final sum of the partial E(T) (replaces world.gop.sum, ccsd_t.h:692): one ncclAllReduce over the per-unit
//     energy vector (x + 0 + ... + 0 is exact, so the result is bit-identical to the single-GPU sum).
// NCCL is bound at run time (dlopen) so the library has no link-time dependency and shares the NCCL a host process
// (MPI/torch) may already have loaded.
#pragma once

#include <dlfcn.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace mpqc_t {

struct NcclUniqueId {
  char internal[128];   // == ncclUniqueId
};
static_assert(sizeof(NcclUniqueId) == sizeof(mpqc_t_unique_id), "mpqc_t_unique_id must be ncclUniqueId-sized");

struct NcclApi {
  typedef struct ncclComm* comm_t;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(comm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};

inline const NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* hnd = nullptr;
    for (const char* n : names) {
      hnd = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (hnd) break;
    }
    if (!hnd) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(hnd, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(hnd, "ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(dlsym(hnd, "ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(hnd, "ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(hnd, "ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(hnd, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(hnd, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.AllGather;
  });
  return api;
}

constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;      // ncclRedOp_t
constexpr int kNcclMin = 3;
constexpr size_t kCommScratchDoubles = size_t(1) << 20;   // 8 MB per device: status words + chunks of the unit-energy vector

inline int nccl_status(int r, const char* what, const char* file, int line) {
  if (r == 0) return MPQC_T_OK;
  const NcclApi& nc = nccl_api();
  char buf[384];
  snprintf(buf, sizeof(buf), "%s -> %s", what, nc.GetErrorString ? nc.GetErrorString(r) : "NCCL error");
  return fail(MPQC_T_ERR_NCCL, buf, file, line);
}
#define MPQC_T_NCCL(expr)                                                             \
  do {                                                                                \
    int _st = ::mpqc_t::nccl_status((expr), #expr, __FILE__, __LINE__);               \
    if (_st != MPQC_T_OK) return _st;                                                 \
  } while (0)

// One member of the communicator that lives in this process: a device, its NCCL communicator, and a small
// pre-allocated device scratch so that the status / energy collectives can never fail for lack of memory.
struct CommMember {
  int rank = 0;              // rank in the (T) communicator
  int device = 0;
  NcclApi::comm_t comm = nullptr;
  double* scratch = nullptr;   // [kCommScratchDoubles]
};

}  // namespace mpqc_t

struct mpqc_t_handle;

struct mpqc_t_comm {
  int nranks = 1;
  bool local = false;                          // true: all nranks members live in this process (one thread per GPU)
  std::vector<mpqc_t::CommMember> members;     // local: nranks entries; rank mode: one
  // Device memory of the last problem, kept per member between calls: allocating and freeing the ~40 GB of operand
  // panels costs seconds per call on some hosts, and a geometry optimisation calls (T) again with the same (o, v).
  // Released by mpqc_t_comm_release_cache / mpqc_t_comm_destroy.
  std::vector<mpqc_t_handle*> cached;
};
