// Shared helpers for libmpqc_t_cuda: error plumbing and the sm_100a PTX wrappers
// (mbarrier, TMA cp.async.bulk.tensor, FP64 tensor-core DMMA).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/mpqc_t.h"

namespace mpqc_t {

// ---------------------------------------------------------------------------------------------
// error plumbing: every failure becomes a status code + a thread-local message; nothing throws
// across the C ABI (SURVEY.md section 8b "Error convention").
// ---------------------------------------------------------------------------------------------
inline std::string& last_error_string() {
  static thread_local std::string s;
  return s;
}

inline int fail(int code, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s (%s:%d)", what, file, line);
  last_error_string() = buf;
  return code;
}

inline int cuda_status(cudaError_t e, const char* expr, const char* file, int line) {
  if (e == cudaSuccess) return MPQC_T_OK;
  char buf[512];
  snprintf(buf, sizeof(buf), "%s -> %s", expr, cudaGetErrorString(e));
  int code = MPQC_T_ERR_CUDA;
  if (e == cudaErrorMemoryAllocation) code = MPQC_T_ERR_OOM;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
    code = MPQC_T_ERR_NO_DEVICE;
  cudaGetLastError();  // clear sticky-less errors
  return fail(code, buf, file, line);
}

#define MPQC_T_CUDA(expr)                                                        \
  do {                                                                           \
    int _st = ::mpqc_t::cuda_status((expr), #expr, __FILE__, __LINE__);          \
    if (_st != MPQC_T_OK) return _st;                                            \
  } while (0)

#define MPQC_T_CHECK(cond, code, msg)                                            \
  do {                                                                           \
    if (!(cond)) return ::mpqc_t::fail((code), (msg), __FILE__, __LINE__);       \
  } while (0)

#define MPQC_T_TRY(expr)                                                         \
  do {                                                                           \
    int _st = (expr);                                                            \
    if (_st != MPQC_T_OK) return _st;                                            \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity);

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef MPQC_T_BOUNDED_CONSUMER_WAIT   // A/B build only (scripts/ab_variants.sh): bounded wait in the consumers too
  mbar_wait_bounded(bar, parity);
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// Bounded wait, used by the TMA producer lane only (the DMMA consumers keep the tight loop above, so their hot path is
// untouched).  A barrier that never completes (a bad tensor map,
// a lost TMA transaction) must surface as an error code at the C ABI, not as a hung GPU: the producer outlives every
// consumer wait (it drains the ring before it exits, w_contract.cuh), so if the pipeline stalls anywhere the producer
// is the thread that notices -- after kMbarTimeoutNs of failed polls it traps, the launch fails with a CUDA error and
// the entry point returns MPQC_T_ERR_CUDA.  Legitimate waits last microseconds.
constexpr unsigned long long kMbarTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 0x3fffu) == 0) {
      const unsigned long long t = global_timer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > kMbarTimeoutNs) __trap();
    }
  }
}

// TMA tiled loads (SASS: UTMALDG).  Coordinates are fastest-dimension first.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4.
// lane holds a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], c0/c1 = C[lane>>2][2*(lane&3) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

}  // namespace mpqc_t
