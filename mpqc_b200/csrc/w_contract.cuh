// W-contraction kernel: the six particle (K = v) and six hole (K = o) contractions of
// ccsd_t.h:314-340 / :359-374 / :1142-1146 at fixed occupied triple (i,j,k), as three
// "pair GEMMs" with a concatenated contraction index kap in [0, v+o):
//
//   N_g[p][q][r] = sum_kap A_x1[p][q][kap] * B_y1z1[r][kap]  +  sum_kap A_x2[q][p][kap] * B_y2z2[r][kap]
//
//   g = 0:  x1 = i, y1z1 = (j,k);  x2 = j, y2z2 = (i,k)       W[a,b,c] += N_0[a][b][c]
//   g = 1:  x1 = i, y1z1 = (k,j);  x2 = k, y2z2 = (i,j)       W[a,b,c] += N_1[a][c][b]
//   g = 2:  x1 = k, y1z1 = (j,i);  x2 = j, y2z2 = (k,i)       W[a,b,c] += N_2[c][b][a]
//
// (the permuted adds happen in the energy kernel's shared-memory tiles, so the reference's five
// 6-index permutes of ccsd_t.h:498-555 never touch HBM).
//
// Machine mapping (sm_100a): persistent grid, one CTA per SM.  One producer warp issues TMA
// (cp.async.bulk.tensor, 128B-swizzled boxes) into a STAGES-deep shared-memory ring guarded by
// full/empty mbarriers; eight consumer warps each own a 16-row x (8*nfrag)-column slice of the
// 128 x tn output tile and issue FP64 tensor-core DMMA.8x8x4 from conflict-free LDS.128 fragment
// reads.  tcgen05/TMEM has no FP64 kind, so DMMA is the Blackwell tensor path for doubles.
//
// Tile rows are a (tp x tq) patch of (p,q); the second term reads the SAME rows from the
// transposed patch A_x2[q][p][:] with a second TMA box, so no transposed copy of A exists.
#pragma once

#include "common.cuh"

namespace mpqc_t {

constexpr int kBM = 128;           // rows of a CTA tile (8 consumer warps x 16)
constexpr int kBK = 16;            // doubles per k-block = one 128-byte swizzled row
constexpr int kMaxNFrag = 16;      // 8-column fragments per tile -> tn <= 128
constexpr int kStages = 5;
constexpr int kConsumerWarps = 8;
constexpr int kGemmThreads = (kConsumerWarps + 4) * 32;   // 2 consumer warpgroups + 1 producer warpgroup
constexpr int kAStageBytes = kBM * kBK * 8;              // 16 KB
constexpr int kBStageBytes = kMaxNFrag * 8 * kBK * 8;    // 16 KB
constexpr int kStageBytes = kAStageBytes + kBStageBytes;
constexpr int kGemmSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct GemmParams {
  int v, o, Kp, kblocks;      // kblocks = ceil(Kp / 16)
  int tp, tq, tn, nfrag;      // row patch, column tile, tn = 8 * nfrag
  int npt, nqt, nnt;          // tile counts along p, q, r
  int tiles_per_group;        // npt * nqt * nnt
  int total_tiles;            // nbatch * 3 * tiles_per_group
  int ldw;                    // row pitch (doubles) of the N_g arrays
  int rows_valid;             // tp * tq
  const int* triples;         // [nbatch][3] (i,j,k) of this launch
  double* w;                  // [nbatch][3][v*v*ldw]
};

// sigma: MMA row/col index g (0..7) -> row inside the 8-row group.  Pairs (2h, 2h+1) map to rows
// (h, h+4) so that the eight lanes of a quarter-warp hit eight distinct 16-byte chunks of the
// 128B-swizzled tile (conflict-free LDS.128).
__device__ __forceinline__ int sigma8(int g) { return (g >> 1) | ((g & 1) << 2); }

__device__ __forceinline__ void decode_tile(const GemmParams& P, int tile, int& b, int& g, int& pt,
                                            int& qt, int& nt) {
  nt = tile % P.nnt;
  int t = tile / P.nnt;
  qt = t % P.nqt;
  t /= P.nqt;
  pt = t % P.npt;
  t /= P.npt;
  g = t % 3;
  b = t / 3;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
w_contract_dmma_kernel(const __grid_constant__ CUtensorMap tmA_n,   // box (16, tq, tp, 1)
                       const __grid_constant__ CUtensorMap tmA_t,   // box (16, tp, tq, 1)
                       const __grid_constant__ CUtensorMap tmB,     // box (16, tn, 1)
                       const GemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const uint32_t a_bytes = (uint32_t)P.rows_valid * kBK * 8;
  const uint32_t b_bytes = (uint32_t)P.tn * kBK * 8;

  if (warp >= kConsumerWarps) {
    // ===================== TMA producer warpgroup (one elected lane works) =====================
    // hand the producer warpgroup's registers to the consumers (each SMSP's 16K-entry file holds
    // two consumer warps + one producer-group warp: 2*232 + 40 <= 512)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == kConsumerWarps && lane == 0) {
      tma_prefetch_desc(&tmA_n);
      tma_prefetch_desc(&tmA_t);
      tma_prefetch_desc(&tmB);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        int b, g, pt, qt, nt;
        decode_tile(P, tile, b, g, pt, qt, nt);
        const int i = P.triples[3 * b + 0], j = P.triples[3 * b + 1], k = P.triples[3 * b + 2];
        int x1, yz1, x2, yz2;
        if (g == 0)      { x1 = i; yz1 = j * P.o + k; x2 = j; yz2 = i * P.o + k; }
        else if (g == 1) { x1 = i; yz1 = k * P.o + j; x2 = k; yz2 = i * P.o + j; }
        else             { x1 = k; yz1 = j * P.o + i; x2 = j; yz2 = k * P.o + i; }
        const int p0 = pt * P.tp, q0 = qt * P.tq, r0 = nt * P.tn;
        for (int term = 0; term < 2; ++term) {
          for (int kb = 0; kb < P.kblocks; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kStageBytes;
            uint8_t* sb = sa + kAStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
            if (term == 0) {
              tma_load_4d(sa, &tmA_n, &full_bar[stage], kb * kBK, q0, p0, x1);
              tma_load_3d(sb, &tmB, &full_bar[stage], kb * kBK, r0, yz1);
            } else {
              tma_load_4d(sa, &tmA_t, &full_bar[stage], kb * kBK, p0, q0, x2);
              tma_load_3d(sb, &tmB, &full_bar[stage], kb * kBK, r0, yz2);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  const int g8 = lane >> 2;          // MMA row (A) / column (B) index
  const int kq = lane & 3;           // MMA k index
  const int sg = sigma8(g8);
  const uint32_t smem_base = smem_u32(smem);

  // B fragment row offsets are tile independent: tile column 8*ni + sigma(g8)
  //   byte offset of chunk c in row r: r*128 + ((c ^ (r & 7)) << 4)
  const uint32_t b_lo_off = (uint32_t)sg * 128u + (uint32_t)((kq ^ sg) << 4);   // + ni * 1024
  // rows of this thread for the normal term: m = 16*warp + 8*mi + sigma(g8)
  uint32_t a_off_n[2], a_off_t[2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    int m = 16 * warp + 8 * mi + sg;
    a_off_n[mi] = (uint32_t)m * 128u + (uint32_t)((kq ^ (m & 7)) << 4);
    int mm = m < P.rows_valid ? m : 0;
    int srow = (mm % P.tq) * P.tp + (mm / P.tq);     // row of (p,q) inside the transposed box
    a_off_t[mi] = (uint32_t)srow * 128u + (uint32_t)((kq ^ (srow & 7)) << 4);
  }
  const int nfrag = P.nfrag;

  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    double acc[2][kMaxNFrag][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < kMaxNFrag; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    for (int term = 0; term < 2; ++term) {
      const uint32_t a0 = term == 0 ? a_off_n[0] : a_off_t[0];
      const uint32_t a1 = term == 0 ? a_off_n[1] : a_off_t[1];
      for (int kb = 0; kb < P.kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint32_t sb = sa + kAStageBytes + b_lo_off;
        const bool second_half = (kb * kBK + 8) < P.Kp;   // kap 8..15 of this block exist
        double2 alo[2], ahi[2];
        alo[0] = lds128(sa + a0);
        alo[1] = lds128(sa + a1);
        ahi[0] = lds128(sa + (a0 ^ 64u));
        ahi[1] = lds128(sa + (a1 ^ 64u));
#pragma unroll
        for (int nc = 0; nc < kMaxNFrag; nc += 4) {
          if (nc < nfrag) {
            double2 blo[4], bhi[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              blo[u] = lds128(sb + (nc + u) * 1024u);
              bhi[u] = lds128((sb + (nc + u) * 1024u) ^ 64u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (nc + u < nfrag) {
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                  dmma884(acc[mi][nc + u][0], acc[mi][nc + u][1], alo[mi].x, blo[u].x);
                  dmma884(acc[mi][nc + u][0], acc[mi][nc + u][1], alo[mi].y, blo[u].y);
                }
                if (second_half) {
#pragma unroll
                  for (int mi = 0; mi < 2; ++mi) {
                    dmma884(acc[mi][nc + u][0], acc[mi][nc + u][1], ahi[mi].x, bhi[u].x);
                    dmma884(acc[mi][nc + u][0], acc[mi][nc + u][1], ahi[mi].y, bhi[u].y);
                  }
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }

    // ---- epilogue: registers -> N_g[p][q][r0 + col] (32-byte sector-aligned runs) ----
    int b, g, pt, qt, nt;
    decode_tile(P, tile, b, g, pt, qt, nt);
    const int p0 = pt * P.tp, q0 = qt * P.tq, r0 = nt * P.tn;
    double* wg = P.w + ((int64_t)(b * 3 + g)) * P.v * P.v * P.ldw;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int m = 16 * warp + 8 * mi + sg;
      const int p = p0 + m / P.tq, q = q0 + m % P.tq;
      const bool row_ok = (m < P.rows_valid) && (p < P.v) && (q < P.v);
      double* row = wg + ((int64_t)p * P.v + q) * P.ldw + r0;
#pragma unroll
      for (int ni = 0; ni < kMaxNFrag; ++ni) {
        if (ni < nfrag && row_ok) {
          // C fragment columns 2*kq, 2*kq+1 of the MMA -> tile columns 8*ni + sigma(2kq), sigma(2kq+1)
          const int c0 = 8 * ni + kq, c1 = 8 * ni + kq + 4;
          if (r0 + c0 < P.v) row[c0] = acc[mi][ni][0];
          if (r0 + c1 < P.v) row[c1] = acc[mi][ni][1];
        }
      }
    }
  }
}

}  // namespace mpqc_t
