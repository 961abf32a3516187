// Host-side helpers of libmpqc_t_cuda: wall clock, cuTensorMapEncodeTiled binding, RAII for CUDA events and device buffers.
#pragma once

#include <chrono>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace mpqc_t;

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn cached = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!cached) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MPQC_T_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    MPQC_T_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, MPQC_T_ERR_CUDA,
                 "cuTensorMapEncodeTiled not available from the driver");
    cached = reinterpret_cast<EncodeTiledFn>(fn);
  }
  *out = cached;
  return MPQC_T_OK;
}

int encode_map(CUtensorMap* map, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
  EncodeTiledFn fn;
  MPQC_T_TRY(get_encode_fn(&fn));
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int d = 0; d < rank; ++d) {
    gdim[d] = dims[d];
    bdim[d] = box[d];
    estr[d] = 1;
    if (d > 0) gstr[d - 1] = strides_bytes[d - 1];
  }
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, base, gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u,%u,%u)", (int)r,
             rank, box[0], box[1], rank > 2 ? box[2] : 0u);
    return fail(MPQC_T_ERR_CUDA, buf, __FILE__, __LINE__);
  }
  return MPQC_T_OK;
}

int64_t roundup(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// CUDA events released on every exit path
struct EventList {
  std::vector<cudaEvent_t> ev;
  int add(cudaEvent_t* out) {
    cudaEvent_t e;
    MPQC_T_CUDA(cudaEventCreate(&e));
    ev.push_back(e);
    *out = e;
    return MPQC_T_OK;
  }
  ~EventList() {
    for (auto e : ev) cudaEventDestroy(e);
  }
};

// device allocation released on every exit path
struct DevBuf {
  double* p = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { cudaFree(p); }
  int alloc(size_t doubles) {
    MPQC_T_CUDA(cudaMalloc(&p, std::max<size_t>(doubles, 1) * sizeof(double)));
    return MPQC_T_OK;
  }
};

}  // namespace
