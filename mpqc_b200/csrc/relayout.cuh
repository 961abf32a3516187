// Integral / amplitude blocking on device: builds the occupied-major, contraction-index-fastest
// operand panels the W-contraction kernel streams with TMA.  Replaces the reference's host-side
// permutes  result("a,i,j,k") = result("i,j,k,a")  etc. (ccsd_t.h:2219,2233,2242) and reblock()
// (ccsd_t.h:2065-2185).  All kernels are HBM-bound tile transposes / strided copies.
//
//   A[x][p][q][kap]   kap <  v : g_abci[kap][p][q][x]      (particle operand, g_dabi of ccsd_t.h:312-314)
//                     kap >= v : -t2[p][q][x][kap-v]       (hole operand, t2_abil of :328-331, sign folded in)
//   B[y][z][r][kap]   kap <  v : t2[kap][r][y][z]          (t2_dcjk of :316-319)
//                     kap >= v : g_aijk[r][y][z][kap-v]    (g_cjkl of :321-326)
//   AT[x][p][q][kap]  = A[x][q][p][kap]  (optional transposed copy: lets the second GEMM term read the same
//                     flattened (p,q) row range as the first; built when 2|A| fits in HBM, "flat" mode)
//   GV[i][j][a][b]    g_abij[a][b][i][j]                   (for the disconnected term, :379-409)
//   T1T[i][a]         t1[a][i]
// kap runs over Kp = roundup8(v+o) >= 16 (zero padded) so that particle and hole terms are ONE
// contraction of length v+o.
#pragma once

#include "common.cuh"

namespace mpqc_t {

// in[kap][mid][j]  ->  out[(j / jdiv) * s1 + (j % jdiv) * s2 + mid * s3 + kap]
__global__ void __launch_bounds__(256)
transpose_kap_last_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t nk,
                          int64_t nmid, int64_t nj, int64_t jdiv, int64_t s1, int64_t s2,
                          int64_t s3) {
  __shared__ double tile[32][33];
  const int64_t mid = blockIdx.z;
  const int64_t j0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int64_t kap = k0 + ty + r, j = j0 + tx;
    double val = 0.0;
    if (kap < nk && j < nj) val = __ldg(in + (kap * nmid + mid) * nj + j);
    tile[ty + r][tx] = val;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    int64_t j = j0 + ty + r, kap = k0 + tx;
    if (kap < nk && j < nj) out[(j / jdiv) * s1 + (j % jdiv) * s2 + mid * s3 + kap] = tile[tx][ty + r];
  }
}

// in[outer][mid][l] (l in [0,o))  ->  out[(outer / odiv) * so1 + (outer % odiv) * so2 + mid * sm + koff + l] = scale * in
__global__ void __launch_bounds__(256)
copy_hole_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n_outer,
                 int64_t n_mid, int64_t o, int64_t odiv, int64_t so1, int64_t so2, int64_t sm, int64_t koff,
                 double scale) {
  const int64_t total = n_outer * n_mid * o;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t l = idx % o;
    int64_t mid = (idx / o) % n_mid;
    int64_t outer = idx / (o * n_mid);
    out[(outer / odiv) * so1 + (outer % odiv) * so2 + mid * sm + koff + l] = scale * __ldg(in + idx);
  }
}

// hole part of ONE operand panel (panel-cache mode): out[row(p,q)][v + l] = -t2[p][q][x][l], with row(p,q) = p*v + q
// (panel A_x) or q*v + p (transposed copy AT_x)
__global__ void __launch_bounds__(256)
copy_hole_panel_kernel(const double* __restrict__ t2, double* __restrict__ panel, int64_t v, int64_t o, int64_t x,
                       int64_t Kp, int transposed) {
  const int64_t total = v * v * o;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t l = idx % o, pq = idx / o;
    const int64_t pp = pq / v, qq = pq % v;
    const int64_t row = transposed ? qq * v + pp : pq;
    panel[row * Kp + v + l] = -__ldg(t2 + (pq * o + x) * o + l);
  }
}

// AT_x[p][q][:] = A_x[q][p][:]: the transposed copy of one complete operand panel (particle AND hole part), whole rows
// of Kp doubles at a time (16-byte vectors; Kp is a multiple of 8).  One warp per row, HBM-bound.
__global__ void __launch_bounds__(256)
transpose_panel_kernel(const double* __restrict__ a, double* __restrict__ at, int64_t v, int64_t Kp) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t nvec = Kp / 2;
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < v * v; row += warps) {
    const int64_t p = row / v, q = row % v;
    const double2* src = reinterpret_cast<const double2*>(a + (q * v + p) * Kp);
    double2* dst = reinterpret_cast<double2*>(at + row * Kp);
    for (int64_t c = lane; c < nvec; c += 32) dst[c] = __ldg(src + c);
  }
}

inline int launch_transpose(cudaStream_t st, const double* in, double* out, int64_t nk, int64_t nmid,
                            int64_t nj, int64_t jdiv, int64_t s1, int64_t s2, int64_t s3,
                            int64_t* launches) {
  // grid.z carries mid and grid.y the kap tiles; split either if it exceeds the 65535 limit
  const int64_t zmax = 65535, kmax = 65535 * 32;
  for (int64_t m0 = 0; m0 < nmid; m0 += zmax) {
    const int64_t mz = nmid - m0 < zmax ? nmid - m0 : zmax;
    for (int64_t k0 = 0; k0 < nk; k0 += kmax) {
      const int64_t kz = nk - k0 < kmax ? nk - k0 : kmax;
      dim3 grid((unsigned)((nj + 31) / 32), (unsigned)((kz + 31) / 32), (unsigned)mz);
      // in[(kap*nmid + mid)*nj + j]: shifting kap by k0 and mid by m0 is a pure pointer offset
      transpose_kap_last_kernel<<<grid, dim3(32, 8), 0, st>>>(in + (k0 * nmid + m0) * nj, out + m0 * s3 + k0,
                                                             kz, nmid, nj, jdiv, s1, s2, s3);
      MPQC_T_CUDA(cudaGetLastError());
      if (launches) ++*launches;
    }
  }
  return MPQC_T_OK;
}

inline int launch_copy_hole(cudaStream_t st, const double* in, double* out, int64_t n_outer,
                            int64_t n_mid, int64_t o, int64_t odiv, int64_t so1, int64_t so2, int64_t sm,
                            int64_t koff, double scale, int64_t* launches) {
  int64_t total = n_outer * n_mid * o;
  if (total == 0) return MPQC_T_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  copy_hole_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, out, n_outer, n_mid, o, odiv, so1, so2, sm, koff, scale);
  MPQC_T_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  return MPQC_T_OK;
}

}  // namespace mpqc_t
