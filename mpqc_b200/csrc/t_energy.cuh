// Fused energy kernel: for one occupied triple (i,j,k) it assembles W from the three pair-GEMM
// outputs, forms the disconnected term V on the fly, applies the (4,1,1,-2,-2,-2) symmetrisation,
// divides by the orbital-energy denominator and reduces.  Replaces, in one pass over W:
//   the five t3 permutes             ccsd_t.h:498-555
//   compute_v3 + its three permutes  ccsd_t.h:379-409, :559-596
//   the symmetrise-multiply          ccsd_t.h:598-609  (:1163-1167 in the straight form)
//   CCSD_T_Reduce / ReduceSymm       ccsd_t.h:2286-2334, :2350-2431
//
//   W[a,b,c]  = N_0[a][b][c] + N_1[a][c][b] + N_2[c][b][a]
//   V[a,b,c]  = g_ij[a,b] t1[c,k] + g_jk[b,c] t1[a,i] + g_ik[a,c] t1[b,j]
//   Z[a,b,c]  = 4 W[abc] + W[bca] + W[cab] - 2 (W[cba] + W[acb] + W[bac])
//   E_ijk     = sum_abc (W+V) Z / (e_i + e_j + e_k - e_a - e_b - e_c)
//
// A block owns one unordered set of three 8-wide virtual tiles {TA >= TB >= TC}.  The 18 raw 8x8x8
// tiles it needs (6 permuted tile coordinates x 3 arrays) are fetched ONCE from HBM/L2 with cp.async
// (16-byte chunks, zero-filled outside the tensor) into XOR-swizzled shared memory, all in flight at
// once.  A thread then owns an element triple (a,b,c) of the (TA,TB,TC) tile and evaluates all six
// permutations of it from shared memory: the six W values are read once, the denominator (symmetric
// in a,b,c) is divided once.  When tiles coincide every element is visited `mult` times (2 or 6), so
// the block sum is scaled by 1/mult.  HBM-bound: algorithmic bytes = 3 arrays * 8 v^3 per triple.
#pragma once

#include "common.cuh"

namespace mpqc_t {

constexpr int kET = 8;              // energy tile edge
constexpr int kEThreads = 512;
constexpr int kETileElems = kET * kET * kET;                 // 512 doubles
constexpr int kEStageBytes = 18 * kETileElems * 8;           // 73,728 B of raw tiles
constexpr int kEGPitch = 9;                                  // padded row pitch of the 8x8 g patches (bank conflicts)
constexpr int kEGPatch = 8 * kEGPitch;                       // doubles per patch
constexpr int kESmallDoubles = 27 * kEGPatch + 72 + 24 + 16;  // g patches, t1 slices, eps slices, reduction
constexpr int kEnergySmemBytes = kEStageBytes + kESmallDoubles * 8;

struct EnergyParams {
  int v, o, ldw;
  int ntile;                  // ceil(v / 8)
  int ntt;                    // ntile (ntile+1) (ntile+2) / 6 tile sets
  const int* triples;         // [nbatch][3]
  const double* w;            // [nbatch][3][v*v*ldw]
  const double* gv;           // [o*o][v][v]
  const double* t1t;          // [o][v]
  const double* eps_occ;      // [o]
  const double* eps_vir;      // [v]
  const uint8_t* tile_sets;   // [ntt][4]  (TA, TB, TC, pad)
  double* partial;            // [nbatch][ntt]
};

// the six permutations s = (s0,s1,s2): tile coordinates (X,Y,Z) = (T[s0], T[s1], T[s2])
#define MPQC_T_PERMS {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}, {0, 2, 1}, {1, 0, 2}}

__host__ __device__ constexpr int perm_index(int s0, int s1) {
  return s0 == 0 ? (s1 == 1 ? 0 : 4) : (s0 == 1 ? (s1 == 2 ? 1 : 5) : (s1 == 0 ? 2 : 3));
}

// swizzled position of element (x,y,z) inside an 8x8x8 tile: rows (x,y) of 8 doubles; adjacent rows
// are swapped and the 16-byte chunk index is XORed so that reads with any one of x,y,z varying fastest
// across the lanes spread over the banks.  Bits 1-2 of z only are permuted, so a (z even, z+1) pair
// stays one contiguous 16-byte chunk (the cp.async granule).
__device__ __forceinline__ int sw_idx(int x, int y, int z) {
  const int row = ((x << 3) | y) ^ (((y >> 2) ^ x) & 1);
  return (row << 3) | (z ^ (((y ^ (x >> 1)) & 3) << 1));
}

__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void cp_async_8_zfill(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(kEThreads, 2)
t_energy_fused_kernel(const EnergyParams P) {
  extern __shared__ __align__(16) uint8_t esmem[];
  double* S = reinterpret_cast<double*>(esmem);                 // [3 arrays][6 perms][512] swizzled
  double* Gs = S + 18 * kETileElems;                            // [3][3][3][8][8]
  double* T1s = Gs + 27 * kEGPatch;                             // [3][3][8]
  double* Ev = T1s + 72;                                        // [3][8]
  double* red = Ev + 24;                                        // [16]

  constexpr int PERM[6][3] = MPQC_T_PERMS;
  const int b = blockIdx.y;
  const int tt = blockIdx.x;
  const int tid = threadIdx.x;
  const int i = P.triples[3 * b], j = P.triples[3 * b + 1], k = P.triples[3 * b + 2];
  const int T[3] = {P.tile_sets[4 * tt], P.tile_sets[4 * tt + 1], P.tile_sets[4 * tt + 2]};
  const int v = P.v, ldw = P.ldw;
  const double* nbase = P.w + (int64_t)(b * 3) * v * v * ldw;
  const int64_t nstride = (int64_t)v * v * ldw;

  // ---- issue all 18 raw-tile copies: thread -> (row = tid>>2, 16-byte chunk = tid&3) of every tile ----
  {
    const int cidx = tid & 255, half = tid >> 8;            // each half-block copies 9 of the 18 tiles
    const int row = cidx >> 2, x = row >> 3, y = row & 7, z = (cidx & 3) << 1;
    const uint32_t s_base = smem_u32(S);
    const uint32_t dst_off = (uint32_t)sw_idx(x, y, z) * 8u;
#pragma unroll
    for (int q = 0; q < 18; ++q) {
      if ((q & 1) == half) {
        const int g = q / 6, pi = q % 6;
        const int X = T[PERM[pi][0]], Y = T[PERM[pi][1]], Z = T[PERM[pi][2]];
        // array 0 tile at (X,Y,Z); array 1 tile at (X,Z,Y); array 2 tile at (Z,Y,X)   [first][second][third]
        const int t0 = g == 2 ? Z : X, t1 = g == 1 ? Z : Y, t2 = g == 0 ? Z : (g == 1 ? Y : X);
        const int g0 = t0 * kET + x, g1 = t1 * kET + y, g2 = t2 * kET + z;
        int nbytes = 0;
        if (g0 < v && g1 < v && g2 < v) nbytes = (g2 + 1 < v) ? 16 : 8;
        const double* src = nbase + g * nstride + ((int64_t)(g0 < v ? g0 : 0) * v + (g1 < v ? g1 : 0)) * ldw +
                            (g2 < v ? g2 : 0);
        cp_async_16_zfill(s_base + (uint32_t)(q * kETileElems * 8) + dst_off, src, nbytes);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- small operands (L2 resident): g_ij / g_jk / g_ik patches, t1 and eps slices; also asynchronous so
  //      that no load latency is serialised in front of the compute phase ----
  {
    const uint32_t g_base = smem_u32(Gs);
#pragma unroll
    for (int it = 0; it < (27 * 64 + kEThreads - 1) / kEThreads; ++it) {
      const int e = it * kEThreads + tid;
      if (e < 27 * 64) {
        const int c = e & 7, r = (e >> 3) & 7, tcol = (e >> 6) % 3, trow = (e / 192) % 3, which = e / 576;
        const int x = which == 1 ? j : i;
        const int y = which == 0 ? j : k;
        const int gr = (trow == 0 ? T[0] : (trow == 1 ? T[1] : T[2])) * kET + r;
        const int gc = (tcol == 0 ? T[0] : (tcol == 1 ? T[1] : T[2])) * kET + c;
        const bool ok = gr < v && gc < v;
        const double* src = P.gv + ((int64_t)(x * P.o + y) * v + (ok ? gr : 0)) * v + (ok ? gc : 0);
        cp_async_8_zfill(g_base + (uint32_t)((e >> 6) * kEGPatch + r * kEGPitch + c) * 8u, src, ok ? 8 : 0);
      }
    }
    if (tid < 72) {
      const int c = tid & 7, t = (tid >> 3) % 3, which = tid / 24;
      const int x = which == 0 ? i : (which == 1 ? j : k);
      const int gc = (t == 0 ? T[0] : (t == 1 ? T[1] : T[2])) * kET + c;
      cp_async_8_zfill(smem_u32(T1s) + (uint32_t)tid * 8u, P.t1t + (int64_t)x * v + (gc < v ? gc : 0), gc < v ? 8 : 0);
    } else if (tid >= 96 && tid < 120) {
      const int c = tid & 7, t = (tid - 96) >> 3;
      const int gc = (t == 0 ? T[0] : (t == 1 ? T[1] : T[2])) * kET + c;
      cp_async_8_zfill(smem_u32(Ev) + (uint32_t)(tid - 96) * 8u, P.eps_vir + (gc < v ? gc : 0), gc < v ? 8 : 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const double eijk = __ldg(P.eps_occ + i) + __ldg(P.eps_occ + j) + __ldg(P.eps_occ + k);

  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- evaluate: thread owns (la,lb,lc) of the (TA,TB,TC) tile and all six permutations of it ----
  double sum = 0.0;
  {
    const int l[3] = {tid >> 6, (tid >> 3) & 7, tid & 7};
    const int ga = T[0] * kET + l[0], gb = T[1] * kET + l[1], gc = T[2] * kET + l[2];
    if (ga < v && gb < v && gc < v) {
      // the 18 reads use only six distinct swizzled offsets: those of the six permutations of (la,lb,lc)
      int swo[6];
#pragma unroll
      for (int pi = 0; pi < 6; ++pi) swo[pi] = sw_idx(l[PERM[pi][0]], l[PERM[pi][1]], l[PERM[pi][2]]);
      double w[6];
#pragma unroll
      for (int pi = 0; pi < 6; ++pi) {
        const int s0 = PERM[pi][0], s1 = PERM[pi][1], s2 = PERM[pi][2];
        // element (e[s0], e[s1], e[s2]) of W = N_0[.s0.][.s1.][.s2.] + N_1[.s0.][.s2.][.s1.] + N_2[.s2.][.s1.][.s0.]
        w[pi] = S[(0 * 6 + pi) * kETileElems + swo[pi]] + S[(1 * 6 + pi) * kETileElems + swo[perm_index(s0, s2)]] +
                S[(2 * 6 + pi) * kETileElems + swo[perm_index(s2, s1)]];
      }
      // Z_s = 4W_s + (two cyclic partners) - 2 (three transposed partners) = 3 W_s + S_same - 2 S_other, where
      // S_even = W[abc]+W[bca]+W[cab] (perms 0,1,2) and S_odd = W[cba]+W[acb]+W[bac] (perms 3,4,5)
      const double s_even = w[0] + w[1] + w[2], s_odd = w[3] + w[4] + w[5];
      double acc = 0.0;
#pragma unroll
      for (int pi = 0; pi < 6; ++pi) {
        const int s0 = PERM[pi][0], s1 = PERM[pi][1], s2 = PERM[pi][2];
        const double z = 3.0 * w[pi] + (pi < 3 ? s_even - 2.0 * s_odd : s_odd - 2.0 * s_even);
        // V for (a',b',c') = (e[s0], e[s1], e[s2]):  g_ij[a',b'] t1[c',k] + g_jk[b',c'] t1[a',i] + g_ik[a',c'] t1[b',j]
        const double vv =
            Gs[((0 * 3 + s0) * 3 + s1) * kEGPatch + l[s0] * kEGPitch + l[s1]] * T1s[(2 * 3 + s2) * 8 + l[s2]] +
            Gs[((1 * 3 + s1) * 3 + s2) * kEGPatch + l[s1] * kEGPitch + l[s2]] * T1s[(0 * 3 + s0) * 8 + l[s0]] +
            Gs[((2 * 3 + s0) * 3 + s2) * kEGPatch + l[s0] * kEGPitch + l[s2]] * T1s[(1 * 3 + s1) * 8 + l[s1]];
        acc += (w[pi] + vv) * z;
      }
      const double d = eijk - Ev[l[0]] - Ev[8 + l[1]] - Ev[16 + l[2]];
      sum += acc / d;
    }
  }
  // every element of the distinct permuted tiles was visited mult times
  const double mult = (T[0] == T[1] && T[1] == T[2]) ? 6.0 : ((T[0] == T[1] || T[1] == T[2]) ? 2.0 : 1.0);
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int wi = 0; wi < kEThreads / 32; ++wi) s += red[wi];
    P.partial[(int64_t)b * P.ntt + tt] = s / mult;
  }
}

// deterministic second stage: unit_e[b] = weight(i,j,k) * sum_tt partial[b][tt] (fixed tree order)
__global__ void __launch_bounds__(256)
t_energy_finish_kernel(const double* __restrict__ partial, int ntt, const int* __restrict__ triples,
                       double* __restrict__ unit_e) {
  __shared__ double red[8];
  const int b = blockIdx.x;
  double s = 0.0;
  for (int t = threadIdx.x; t < ntt; t += 256) s += partial[(int64_t)b * ntt + t];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int wi = 0; wi < 8; ++wi) tot += red[wi];
    const int i = triples[3 * b], j = triples[3 * b + 1], k = triples[3 * b + 2];
    // weights of CCSD_T_ReduceSymm (ccsd_t.h:2399-2423) applied to the occupied triple
    double wgt = (i == j && j == k) ? 0.0 : ((i == j || j == k || i == k) ? 1.0 : 2.0);
    unit_e[b] = wgt * tot;
  }
}

// Decomposition of the same energy over virtual-block triples (validation aid, mpqc_t_run_vblocks):
//   vblock_e[tt] += sum_b weight(i,j,k)_b * partial[b][tt]
// tt enumerates 8-wide virtual tiles TA >= TB >= TC exactly like global_iter - 1 of the reference's coarse loop with
// block size 8 (ccsd_t.h:443-480), so vblock_e[tt] equals the energy that loop iteration adds (ccsd_t.h:619-638).
// One thread per tt walks the batch in order: deterministic.
__global__ void __launch_bounds__(256)
t_energy_vblock_kernel(const double* __restrict__ partial, int ntt, int nbatch, const int* __restrict__ triples,
                       double* __restrict__ vblock_e) {
  const int tt = blockIdx.x * blockDim.x + threadIdx.x;
  if (tt >= ntt) return;
  double s = 0.0;
  for (int b = 0; b < nbatch; ++b) {
    const int i = triples[3 * b], j = triples[3 * b + 1], k = triples[3 * b + 2];
    const double wgt = (i == j && j == k) ? 0.0 : ((i == j || j == k || i == k) ? 1.0 : 2.0);
    s += wgt * partial[(int64_t)b * ntt + tt];
  }
  vblock_e[tt] += s;
}

// W^{abc}_{ijk} itself as a dense [v][v][v] array (mpqc_t_w_batch: the hook iterative-triples models such as CC3 /
// CCSDT-1 would call every iteration, cc3.h:55+, ccsdt1.h:55+):  W[a][b][c] = N_0[a][b][c] + N_1[a][c][b] + N_2[c][b][a].
// One 512-thread block per 8x8x8 output tile: the three source tiles (at permuted tile coordinates) are read with
// 64-byte row segments into shared memory and the permuted adds happen there, so every global access is a row.
__global__ void __launch_bounds__(512)
w_assemble_kernel(const double* __restrict__ n_all, double* __restrict__ w_out, int v, int ldw, int ntile) {
  __shared__ double s[3][8][8][9];
  const int b = blockIdx.y;
  int t = blockIdx.x;
  const int TC = t % ntile;
  t /= ntile;
  const int TB = t % ntile, TA = t / ntile;
  const int64_t nstride = (int64_t)v * v * ldw;
  const double* n0 = n_all + (int64_t)b * 3 * nstride;
  const int x = threadIdx.x >> 6, y = (threadIdx.x >> 3) & 7, z = threadIdx.x & 7;
  // tile (X,Y,Z) of an array: element (x,y,z) at [(X*8+x)*v + (Y*8+y)]*ldw + Z*8+z
  auto fetch = [&](const double* base, int X, int Y, int Z) -> double {
    const int g0 = X * 8 + x, g1 = Y * 8 + y, g2 = Z * 8 + z;
    return (g0 < v && g1 < v && g2 < v) ? __ldg(base + ((int64_t)g0 * v + g1) * ldw + g2) : 0.0;
  };
  s[0][x][y][z] = fetch(n0, TA, TB, TC);                  // N_0[a][b][c]
  s[1][x][y][z] = fetch(n0 + nstride, TA, TC, TB);        // N_1[a][c][b]   (x,y,z) = (a,c,b)
  s[2][x][y][z] = fetch(n0 + 2 * nstride, TC, TB, TA);    // N_2[c][b][a]   (x,y,z) = (c,b,a)
  __syncthreads();
  const int a = TA * 8 + x, bb = TB * 8 + y, c = TC * 8 + z;
  if (a < v && bb < v && c < v)
    w_out[(int64_t)b * v * v * v + ((int64_t)a * v + bb) * v + c] = s[0][x][y][z] + s[1][x][z][y] + s[2][z][y][x];
}

}  // namespace mpqc_t
