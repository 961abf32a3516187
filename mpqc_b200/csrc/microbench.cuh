// FP64 pipe microbenchmarks: fix the roofline denominator for the W contraction on the box
// (MEASURED_PEAKS.json has no FP64 entry).  Issue-bound register-only loops, no memory traffic.
#pragma once

#include "common.cuh"

namespace mpqc_t {

__global__ void __launch_bounds__(256) microbench_dmma_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) c[u][0] = c[u][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) dmma884(c[u][0], c[u][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += c[u][0] + c[u][1];
  if (s == 123.456) out[0] = s;   // keep the loop alive
}

__global__ void __launch_bounds__(256) microbench_dfma_kernel(double* out, int iters) {
  double c[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) c[u] = threadIdx.x * 1e-3 + u;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = fma(c[u], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) s += c[u];
  if (s == 123.456) out[0] = s;
}

}  // namespace mpqc_t
