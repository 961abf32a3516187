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
// A block owns one unordered set of three 8-wide virtual tiles {TA >= TB >= TC}: the six permuted
// 8x8x8 tiles of each N_g are read exactly once from HBM/L2 (64-byte row segments), transposed
// in shared memory, and every (a,b,c) of the distinct permuted tiles is evaluated from shared
// memory.  HBM-bound: algorithmic bytes = 3 arrays * 8 v^3 per triple.
#pragma once

#include "common.cuh"

namespace mpqc_t {

constexpr int kET = 8;              // energy tile edge
constexpr int kEThreads = 256;

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
__constant__ int8_t c_perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}, {0, 2, 1}, {1, 0, 2}};

__device__ __forceinline__ int perm_index(int s0, int s1, int s2) {
  // inverse of c_perm
  if (s0 == 0) return s1 == 1 ? 0 : 4;
  if (s0 == 1) return s1 == 2 ? 1 : 5;
  return s1 == 0 ? 2 : 3;
}

__global__ void __launch_bounds__(kEThreads)
t_energy_fused_kernel(const EnergyParams P) {
  // W tiles of the six permuted tile coordinates; [perm][a][b][c] with a padded fastest pitch
  __shared__ double Wt[6][kET][kET][kET + 1];
  __shared__ double Gs[3][3][3][kET][kET];   // [ij|jk|ik][row tile][col tile][.][.]
  __shared__ double T1s[3][3][kET];          // [i|j|k][tile][.]
  __shared__ double Ev[3][kET];
  __shared__ double red[kEThreads / 32];

  const int b = blockIdx.y;
  const int tt = blockIdx.x;
  const int tid = threadIdx.x;
  const int i = P.triples[3 * b], j = P.triples[3 * b + 1], k = P.triples[3 * b + 2];
  int T[3] = {P.tile_sets[4 * tt], P.tile_sets[4 * tt + 1], P.tile_sets[4 * tt + 2]};
  const int v = P.v, ldw = P.ldw;
  const double* n0 = P.w + (int64_t)(b * 3) * v * v * ldw;
  const double* n1 = n0 + (int64_t)v * v * ldw;
  const double* n2 = n1 + (int64_t)v * v * ldw;

  // ---- small operands: g_ij / g_jk / g_ik patches, t1 and eps slices ----
  for (int e = tid; e < 3 * 3 * 3 * 64; e += kEThreads) {
    int c = e & 7, r = (e >> 3) & 7, tcol = (e >> 6) % 3, trow = (e / 192) % 3, which = e / 576;
    int x = which == 0 ? i : (which == 1 ? j : i);
    int y = which == 0 ? j : k;
    int gr = T[trow] * kET + r, gc = T[tcol] * kET + c;
    double val = 0.0;
    if (gr < v && gc < v) val = __ldg(P.gv + ((int64_t)(x * P.o + y) * v + gr) * v + gc);
    Gs[which][trow][tcol][r][c] = val;
  }
  if (tid < 72) {
    int c = tid & 7, t = (tid >> 3) % 3, which = tid / 24;
    int x = which == 0 ? i : (which == 1 ? j : k);
    int gc = T[t] * kET + c;
    T1s[which][t][c] = gc < v ? __ldg(P.t1t + (int64_t)x * v + gc) : 0.0;
  } else if (tid >= 96 && tid < 120) {
    int c = tid & 7, t = (tid - 96) >> 3;
    int gc = T[t] * kET + c;
    Ev[t][c] = gc < v ? __ldg(P.eps_vir + gc) : 0.0;
  }

  // ---- phase 1: Wt[pi][a][b][c] = N_0[a][b][c]   (rows (a,b), c contiguous) ----
  // 6 perms * 64 rows * 4 double2 = 1536 vector loads, 6 per thread
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    int e = it * kEThreads + tid;
    int c2 = (e & 3) * 2, row = (e >> 2) & 63, pi = e >> 8;
    int la = row >> 3, lb = row & 7;
    int X = T[c_perm[pi][0]], Y = T[c_perm[pi][1]], Z = T[c_perm[pi][2]];
    int ga = X * kET + la, gb = Y * kET + lb, gc = Z * kET + c2;
    double2 val = make_double2(0.0, 0.0);
    if (ga < v && gb < v && gc < v) {
      const double* src = n0 + ((int64_t)ga * v + gb) * ldw + gc;
      val = __ldg(reinterpret_cast<const double2*>(src));   // ldw even, gc even -> 16B aligned
      if (gc + 1 >= v) val.y = 0.0;
    }
    Wt[pi][la][lb][c2] = val.x;
    Wt[pi][la][lb][c2 + 1] = val.y;
  }
  __syncthreads();
  // ---- phase 2: Wt[pi][a][b][c] += N_1[a][c][b]   (rows (a,c), b contiguous) ----
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    int e = it * kEThreads + tid;
    int b2 = (e & 3) * 2, row = (e >> 2) & 63, pi = e >> 8;
    int la = row >> 3, lc = row & 7;
    int X = T[c_perm[pi][0]], Y = T[c_perm[pi][1]], Z = T[c_perm[pi][2]];
    int ga = X * kET + la, gc = Z * kET + lc, gb = Y * kET + b2;
    double2 val = make_double2(0.0, 0.0);
    if (ga < v && gc < v && gb < v) {
      const double* src = n1 + ((int64_t)ga * v + gc) * ldw + gb;
      val = __ldg(reinterpret_cast<const double2*>(src));
      if (gb + 1 >= v) val.y = 0.0;
    }
    Wt[pi][la][b2][lc] += val.x;
    Wt[pi][la][b2 + 1][lc] += val.y;
  }
  __syncthreads();
  // ---- phase 3: Wt[pi][a][b][c] += N_2[c][b][a]   (rows (c,b), a contiguous) ----
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    int e = it * kEThreads + tid;
    int a2 = (e & 3) * 2, row = (e >> 2) & 63, pi = e >> 8;
    int lc = row >> 3, lb = row & 7;
    int X = T[c_perm[pi][0]], Y = T[c_perm[pi][1]], Z = T[c_perm[pi][2]];
    int gc = Z * kET + lc, gb = Y * kET + lb, ga = X * kET + a2;
    double2 val = make_double2(0.0, 0.0);
    if (gc < v && gb < v && ga < v) {
      const double* src = n2 + ((int64_t)gc * v + gb) * ldw + ga;
      val = __ldg(reinterpret_cast<const double2*>(src));
      if (ga + 1 >= v) val.y = 0.0;
    }
    Wt[pi][a2][lb][lc] += val.x;
    Wt[pi][a2 + 1][lb][lc] += val.y;
  }
  __syncthreads();

  // ---- evaluate every element of the distinct permuted tiles ----
  const double eijk = __ldg(P.eps_occ + i) + __ldg(P.eps_occ + j) + __ldg(P.eps_occ + k);
  double sum = 0.0;
#pragma unroll 1
  for (int pi = 0; pi < 6; ++pi) {
    const int s0 = c_perm[pi][0], s1 = c_perm[pi][1], s2 = c_perm[pi][2];
    // skip a permutation whose tile coordinates repeat an earlier one (TA==TB and/or TB==TC)
    bool dup = false;
    for (int pj = 0; pj < pi; ++pj)
      dup |= (T[c_perm[pj][0]] == T[s0]) && (T[c_perm[pj][1]] == T[s1]) && (T[c_perm[pj][2]] == T[s2]);
    if (dup) continue;
    // element permutations: W[b,c,a] lives in the tile with coordinates (Y,Z,X) = perm (s1,s2,s0), ...
    const int p_bca = perm_index(s1, s2, s0), p_cab = perm_index(s2, s0, s1);
    const int p_cba = perm_index(s2, s1, s0), p_acb = perm_index(s0, s2, s1);
    const int p_bac = perm_index(s1, s0, s2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int e = h * kEThreads + tid;
      const int lc = e & 7, lb = (e >> 3) & 7, la = e >> 6;
      const int ga = T[s0] * kET + la, gb = T[s1] * kET + lb, gc = T[s2] * kET + lc;
      if (ga < v && gb < v && gc < v) {
        const double w = Wt[pi][la][lb][lc];
        const double z = 4.0 * w + Wt[p_bca][lb][lc][la] + Wt[p_cab][lc][la][lb] -
                         2.0 * (Wt[p_cba][lc][lb][la] + Wt[p_acb][la][lc][lb] + Wt[p_bac][lb][la][lc]);
        const double vv = Gs[0][s0][s1][la][lb] * T1s[2][s2][lc] + Gs[1][s1][s2][lb][lc] * T1s[0][s0][la] +
                          Gs[2][s0][s2][la][lc] * T1s[1][s1][lb];
        const double d = eijk - Ev[s0][la] - Ev[s1][lb] - Ev[s2][lc];
        sum += (w + vv) * z / d;
      }
    }
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
#pragma unroll
    for (int wi = 0; wi < kEThreads / 32; ++wi) s += red[wi];
    P.partial[(int64_t)b * P.ntt + tt] = s;
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

}  // namespace mpqc_t
