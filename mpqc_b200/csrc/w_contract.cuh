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
// Machine mapping (sm_100a): persistent grid, one CTA per SM, three warpgroups.  One producer lane
// issues TMA (cp.async.bulk.tensor, 128B-swizzled boxes) into a kStages-deep shared-memory ring (two k-blocks per stage) guarded
// by full/empty mbarriers; eight consumer warps each own a 16-row x (8*NFRAG)-column slice of the
// 128 x tn output tile and issue FP64 tensor-core DMMA.8x8x4 from conflict-free LDS.128 fragment reads.
// tcgen05/TMEM has no FP64 kind, so DMMA is the Blackwell tensor path for doubles.
//
// Consumer schedule (what the r01a ncu profile asked for): NFRAG is a template parameter so the k-block
// body is branch-free; B fragments are processed in register chunks of <= 4 column fragments, double
// buffered, and the first chunk + A fragments of the NEXT k-block are fetched (after its full-barrier
// wait) before the last chunk of the current block is issued, so no LDS latency is exposed at block
// boundaries; within a chunk DMMAs are issued k-step-major so dependent DMMAs on one accumulator are
// separated by 6-8 independent ones; the half-filled last k-block (Kp % 16 == 8) is a separate
// instantiation instead of predicated-off DMMAs.
//
// Tile rows, two modes.  "flat" (default when 2|A| fits in HBM): 128 consecutive flattened (p,q) rows; the
// second term reads the same row range from the transposed copy AT[x][p][q][:] = A[x][q][p][:], so there is
// no row padding at all.  "patch" (large problems): rows are a (tp x tq) patch of (p,q) and the second term
// reads the SAME rows from the transposed patch A_x2[q][p][:] with a second TMA box of the same memory.
#pragma once

#include "common.cuh"

namespace mpqc_t {

constexpr int kBM = 128;           // rows of a CTA tile (8 consumer warps x 16)
constexpr int kBK = 16;            // doubles per k-block = one 128-byte swizzled row
constexpr int kMaxNFrag = 16;      // 8-column fragments per tile -> tn <= 128
constexpr int kSub = 2;             // k-blocks per pipeline stage: one full/empty barrier round trip per 32 kap
#ifndef MPQC_T_STAGES
#define MPQC_T_STAGES 3
#endif
constexpr int kStages = MPQC_T_STAGES;
constexpr int kConsumerWarps = 8;
constexpr int kGemmThreads = (kConsumerWarps + 4) * 32;   // 2 consumer warpgroups + 1 producer warpgroup
constexpr int kAStageBytes = kBM * kBK * 8;              // 16 KB  (A rows of one k-block)
constexpr int kBStageBytes = kMaxNFrag * 8 * kBK * 8;    // 16 KB  (B rows of one k-block)
constexpr int kSubBytes = kAStageBytes + kBStageBytes;   // one k-block
constexpr int kStageBytes = kSub * kSubBytes;            // 64 KB
constexpr int kGemmSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct GemmParams {
  int v, o, Kp, kblocks;      // kblocks = ceil(Kp / 16)
  int tp, tq, tn, nfrag;      // row patch (patch mode), column tile, tn = 8 * nfrag
  int npt, nqt, nnt;          // tile counts along p, q (patch mode), r
  int skip_last;              // 1: the last column tile drops its trailing fragment (tn*nnt - 8 >= v)
  int flat;                   // 1: rows are 128 consecutive flattened (p,q) and term 1 reads the transposed copy AT
  int nmt;                    // row tiles: ceil(v*v/128) (flat) or npt*nqt (patch)
  int tiles_per_group;        // nmt * nnt
  int total_tiles;            // nbatch * 3 * tiles_per_group
  int main_tiles;             // tiles computing all NFRAG fragments; the remaining ones drop the last fragment
  int ldw;                    // row pitch (doubles) of the N_g arrays
  int rows_valid;             // tp * tq
  const int* triples;         // [nbatch][3] (i,j,k) of this launch
  double* w;                  // [nbatch][3][v*v*ldw]
  const int* a_slot;          // panel-cache mode: occupied index x -> slot of A_x / AT_x in the panel pool (NULL: slot = x)
  // ---- mode 1: plain batched NT GEMM on the same pipeline,  C_b[m][n] = sum_kap L_lb[m][kap] * R_rb[n][kap]
  //      (used to assemble the integral classes from the three-centre factors, df_build.cuh).  tmA_n maps L as
  //      (kap, m, batch) with a (16, 128, 1) box, tmB maps R as (kap, n, batch) with a (16, tn, 1) box;
  //      lb = (b / l_div) % l_mod, rb = (b / r_div) % r_mod; C_b = w + (b / o_div) * out_s1 + (b % o_div) * out_s2,
  //      row pitch ldw64, valid rows m < v, valid columns n < ncols.
  int mode;                   // 0: W contraction (two terms, three groups), 1: plain NT GEMM (one term)
  int ncols;                  // valid output columns (mode 0: v)
  int l_div, l_mod, r_div, r_mod, o_div;
  long long out_s1, out_s2, ldw64;
};

// sigma: MMA row/col index g (0..7) -> row inside the 8-row group.  Pairs (2h, 2h+1) map to rows
// (h, h+4) so that the eight lanes of a quarter-warp hit eight distinct 16-byte chunks of the
// 128B-swizzled tile (conflict-free LDS.128).
__device__ __forceinline__ int sigma8(int g) { return (g >> 1) | ((g & 1) << 2); }

// Tile order: first every "main" tile (column tiles that compute all NFRAG fragments), then the "skip" tiles
// (the last column tile of each row of tiles when it drops its trailing fragment), so that a CTA walking
// tile, tile + grid, ... runs one instantiation of the k-loop, then the other, and the short tiles fill the tail.
__device__ __forceinline__ void decode_tile(const GemmParams& P, int tile, int& b, int& g, int& mt, int& nt) {
  int t;
  if (tile < P.main_tiles) {
    const int nn = P.nnt - P.skip_last;
    nt = tile % nn;
    t = tile / nn;
  } else {
    nt = P.nnt - 1;
    t = tile - P.main_tiles;
  }
  // group fastest, then row tile: the CTAs of one wave that share a row tile read the same operand panels
  // (A_i: g = 0,1 term 0;  AT_j: g = 0,2 term 1;  A_k / AT_k: g = 2 / g = 1) while they are still in L2
  g = t % 3;
  t /= 3;
  mt = t % P.nmt;
  b = t / P.nmt;
}

// mode 1: column tile fastest (main tiles first, like above), then row tile, then batch entry
__device__ __forceinline__ void decode_tile_plain(const GemmParams& P, int tile, int& b, int& mt, int& nt) {
  int t;
  if (tile < P.main_tiles) {
    const int nn = P.nnt - P.skip_last;
    nt = tile % nn;
    t = tile / nn;
  } else {
    nt = P.nnt - 1;
    t = tile - P.main_tiles;
  }
  mt = t % P.nmt;
  b = t / P.nmt;
}

// compile-time chunking of the NFRAG column fragments into register chunks of nearly equal size
template <int NFRAG>
struct Chunking {
  static constexpr int kMax = NFRAG <= 13 ? 4 : 3;              // fragments per chunk (register budget)
  static constexpr int kNum = (NFRAG + kMax - 1) / kMax;        // chunks per k-block
  __host__ __device__ static constexpr int beg(int c) { return c * NFRAG / kNum; }
  __host__ __device__ static constexpr int len(int c) { return beg(c + 1) - beg(c); }
};

struct BFrag {
  double2 lo, hi;   // kap (2k, 2k+1) and (8+2k, 9+2k) of this lane's B row
};

template <int NFRAG, int C>
__device__ __forceinline__ void load_b_chunk(BFrag (&b)[Chunking<NFRAG>::kMax], uint32_t sb) {
  using CH = Chunking<NFRAG>;
#pragma unroll
  for (int u = 0; u < CH::len(C); ++u) {
    const uint32_t addr = sb + (uint32_t)(CH::beg(C) + u) * 1024u;
    b[u].lo = lds128(addr);
    b[u].hi = lds128(addr ^ 64u);
  }
}

// SKIP (0/1): number of trailing column fragments of the LAST chunk that this tile does not compute (the last
// column tile of a row of tiles when 8*NFRAG*nnt overshoots v by a whole fragment); compile-time, so no
// predicated DMMAs.
template <int NFRAG, int C, bool HALF, int SKIP>
__device__ __forceinline__ void mma_chunk(double (&acc)[2][NFRAG][2], const double2 (&alo)[2], const double2 (&ahi)[2],
                                          const BFrag (&b)[Chunking<NFRAG>::kMax]) {
  using CH = Chunking<NFRAG>;
  constexpr int n0 = CH::beg(C), nl = CH::len(C) - ((C == CH::kNum - 1) ? SKIP : 0);
  // k-step major: consecutive DMMAs touch different accumulators
#pragma unroll
  for (int u = 0; u < nl; ++u)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][n0 + u][0], acc[mi][n0 + u][1], alo[mi].x, b[u].lo.x);
#pragma unroll
  for (int u = 0; u < nl; ++u)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][n0 + u][0], acc[mi][n0 + u][1], alo[mi].y, b[u].lo.y);
  if (!HALF) {
#pragma unroll
    for (int u = 0; u < nl; ++u)
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][n0 + u][0], acc[mi][n0 + u][1], ahi[mi].x, b[u].hi.x);
#pragma unroll
    for (int u = 0; u < nl; ++u)
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][n0 + u][0], acc[mi][n0 + u][1], ahi[mi].y, b[u].hi.y);
  }
}

// state a consumer warp carries across k-blocks
template <int NFRAG>
struct ConsumerRegs {
  double2 alo[2], ahi[2];                         // A fragments of the current k-block
  BFrag bb[2][Chunking<NFRAG>::kMax];             // double-buffered B chunks
};

// One k-block.  On entry R holds the A fragments and B chunk 0 (in bb[0]) of this block.  On exit, if
// has_next, it holds those of the next block (read from its stage slot, after waiting on that stage's full
// barrier when the next block is the first of a new stage).
template <int NFRAG, bool HALF, int SKIP>
__device__ __forceinline__ void kblock(double (&acc)[2][NFRAG][2], ConsumerRegs<NFRAG>& R, uint32_t sb_cur,
                                       bool has_next, bool need_wait, uint64_t* next_full, uint32_t next_phase,
                                       uint32_t sa_next, uint32_t a0_next, uint32_t a1_next, uint32_t sb_next) {
  using CH = Chunking<NFRAG>;
  double2 nlo[2], nhi[2];
#pragma unroll
  for (int c = 0; c < CH::kNum; ++c) {
    if (c + 1 < CH::kNum) {
      // prefetch the next chunk of this block (template index must be a constant: unrolled switch)
      if (c == 0) load_b_chunk<NFRAG, (1 < CH::kNum ? 1 : 0)>(R.bb[1], sb_cur);
      if (c == 1) load_b_chunk<NFRAG, (2 < CH::kNum ? 2 : 0)>(R.bb[0], sb_cur);
      if (c == 2) load_b_chunk<NFRAG, (3 < CH::kNum ? 3 : 0)>(R.bb[1], sb_cur);
      if (c == 3) load_b_chunk<NFRAG, (4 < CH::kNum ? 4 : 0)>(R.bb[0], sb_cur);
      if (c == 4) load_b_chunk<NFRAG, (5 < CH::kNum ? 5 : 0)>(R.bb[1], sb_cur);
    } else if (has_next) {
      // last chunk: fetch the next k-block's A fragments and B chunk 0 before issuing this chunk's DMMAs
      if (need_wait) mbar_wait(next_full, next_phase);   // the next k-block opens a new pipeline stage
      nlo[0] = lds128(sa_next + a0_next);
      nlo[1] = lds128(sa_next + a1_next);
      nhi[0] = lds128(sa_next + (a0_next ^ 64u));
      nhi[1] = lds128(sa_next + (a1_next ^ 64u));
      load_b_chunk<NFRAG, 0>(R.bb[(c + 1) & 1], sb_next);
    }
    if (c == 0) mma_chunk<NFRAG, 0, HALF, SKIP>(acc, R.alo, R.ahi, R.bb[0]);
    if (c == 1) mma_chunk<NFRAG, (1 < CH::kNum ? 1 : 0), HALF, SKIP>(acc, R.alo, R.ahi, R.bb[1]);
    if (c == 2) mma_chunk<NFRAG, (2 < CH::kNum ? 2 : 0), HALF, SKIP>(acc, R.alo, R.ahi, R.bb[0]);
    if (c == 3) mma_chunk<NFRAG, (3 < CH::kNum ? 3 : 0), HALF, SKIP>(acc, R.alo, R.ahi, R.bb[1]);
    if (c == 4) mma_chunk<NFRAG, (4 < CH::kNum ? 4 : 0), HALF, SKIP>(acc, R.alo, R.ahi, R.bb[0]);
    if (c == 5) mma_chunk<NFRAG, (5 < CH::kNum ? 5 : 0), HALF, SKIP>(acc, R.alo, R.ahi, R.bb[1]);
  }
  if (has_next) {
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      R.alo[mi] = nlo[mi];
      R.ahi[mi] = nhi[mi];
    }
    if (CH::kNum & 1) {   // chunk 0 of the next block was loaded into bb[1]; the next block expects bb[0]
#pragma unroll
      for (int u = 0; u < CH::kMax; ++u) R.bb[0][u] = R.bb[1][u];
    }
  }
}

// TMA producer (one elected lane of the producer warpgroup): walks this CTA's tiles and streams, per pipeline stage,
// kSub k-blocks of the A rows and the B rows into the shared-memory ring.
__device__ __forceinline__ void producer_loop(const GemmParams& P, const CUtensorMap* tmA_n, const CUtensorMap* tmA_t,
                                              const CUtensorMap* tmB, uint8_t* smem, uint64_t* full_bar,
                                              uint64_t* empty_bar, const uint32_t a_bytes, const uint32_t b_bytes) {
      tma_prefetch_desc(tmA_n);
      tma_prefetch_desc(tmA_t);
      tma_prefetch_desc(tmB);
      int stage = 0;
      uint32_t phase = 0;
      const int nterms = P.mode == 0 ? 2 : 1;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        int b, g, mt, nt;
        int x1, yz1, x2 = 0, yz2 = 0;
        if (P.mode == 0) {
          decode_tile(P, tile, b, g, mt, nt);
          const int i = P.triples[3 * b + 0], j = P.triples[3 * b + 1], k = P.triples[3 * b + 2];
          if (g == 0)      { x1 = i; yz1 = j * P.o + k; x2 = j; yz2 = i * P.o + k; }
          else if (g == 1) { x1 = i; yz1 = k * P.o + j; x2 = k; yz2 = i * P.o + j; }
          else             { x1 = k; yz1 = j * P.o + i; x2 = j; yz2 = k * P.o + i; }
          if (P.a_slot) { x1 = P.a_slot[x1]; x2 = P.a_slot[x2]; }
        } else {
          decode_tile_plain(P, tile, b, mt, nt);
          x1 = (b / P.l_div) % P.l_mod;
          yz1 = (b / P.r_div) % P.r_mod;
        }
        const int p0 = (mt / P.nqt) * P.tp, q0 = (mt % P.nqt) * P.tq, r0 = nt * P.tn, m0 = mt * kBM;
        for (int term = 0; term < nterms; ++term) {
          for (int kb = 0; kb < P.kblocks; kb += kSub) {
            const int nsub = (P.kblocks - kb) < kSub ? (P.kblocks - kb) : kSub;
            mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)nsub * (a_bytes + b_bytes));
            for (int sub = 0; sub < nsub; ++sub) {
              uint8_t* sa = smem + stage * kStageBytes + sub * kSubBytes;
              uint8_t* sb = sa + kAStageBytes;
              const int kap0 = (kb + sub) * kBK;
              if (term == 0) {
                if (P.flat) tma_load_3d(sa, tmA_n, &full_bar[stage], kap0, m0, x1);
                else        tma_load_4d(sa, tmA_n, &full_bar[stage], kap0, q0, p0, x1);
                tma_load_3d(sb, tmB, &full_bar[stage], kap0, r0, yz1);
              } else {
                if (P.flat) tma_load_3d(sa, tmA_t, &full_bar[stage], kap0, m0, x2);
                else        tma_load_4d(sa, tmA_t, &full_bar[stage], kap0, p0, q0, x2);
                tma_load_3d(sb, tmB, &full_bar[stage], kap0, r0, yz2);
              }
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
      // drain: stay until the consumers have released every slot of the last fills, so that this (bounded) lane
      // outlives all their (unbounded) waits -- a stalled pipeline then ends in a trap, not in a hang
      for (int s = 0; s < kStages; ++s) {
        mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
}

// Consumer tile loop for tiles [tile, tile_end) of this CTA (stride gridDim.x); SKIP as in mma_chunk.
template <int NFRAG, int SKIP>
__device__ __forceinline__ void consume_tiles(const GemmParams& P, int& tile, const int tile_end, ConsumerRegs<NFRAG>& R,
                                              int& stage, uint32_t& phase, int& sub, uint64_t* full_bar, uint64_t* empty_bar,
                                              const uint32_t smem_base, const uint32_t b_lo_off,
                                              const uint32_t (&a_off_n)[2], const uint32_t (&a_off_t)[2], const int warp,
                                              const int lane, const int sg, const int kq) {
  const int kblocks = P.kblocks;
  const int nblk = (P.mode == 0 ? 2 : 1) * kblocks;   // k-blocks per tile (two terms in the W contraction)
  const bool half_last = (P.Kp % kBK) != 0;            // Kp % 16 == 8: last block of a term is half filled
  for (; tile < tile_end; tile += gridDim.x) {
    double acc[2][NFRAG][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < NFRAG; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const bool more_tiles = tile + (int)gridDim.x < P.total_tiles;

    for (int q = 0; q < nblk; ++q) {
      const bool last_in_tile = (q == nblk - 1);
      const bool last_in_term = (q == kblocks - 1) || last_in_tile;
      const bool has_next = !last_in_tile || more_tiles;
      const bool next_transposed = !last_in_tile && (q + 1 >= kblocks);
      // the next k-block lives in the other half of this stage, or opens the next stage
      const bool new_stage = last_in_term || (sub == kSub - 1);
      const int ns = new_stage ? ((stage + 1 == kStages) ? 0 : stage + 1) : stage;
      const uint32_t nph = (new_stage && ns == 0) ? (phase ^ 1u) : phase;
      const int nsub = new_stage ? 0 : sub + 1;
      const uint32_t sa_next = smem_base + ns * kStageBytes + nsub * kSubBytes;
      const uint32_t a0n = next_transposed ? a_off_t[0] : a_off_n[0];
      const uint32_t a1n = next_transposed ? a_off_t[1] : a_off_n[1];
      const uint32_t sb_cur = smem_base + stage * kStageBytes + sub * kSubBytes + b_lo_off;
      const bool half = half_last && (q == kblocks - 1 || q == nblk - 1);
      if (half)
        kblock<NFRAG, true, SKIP>(acc, R, sb_cur, has_next, new_stage, &full_bar[ns], nph, sa_next, a0n, a1n,
                                  sa_next + b_lo_off);
      else
        kblock<NFRAG, false, SKIP>(acc, R, sb_cur, has_next, new_stage, &full_bar[ns], nph, sa_next, a0n, a1n,
                                   sa_next + b_lo_off);
      if (new_stage) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
      }
      stage = ns;
      phase = nph;
      sub = nsub;
    }

    // ---- epilogue: registers -> N_g[p][q][r0 + col] (32-byte sector-aligned runs) ----
    int b, g = 0, mt, nt;
    if (P.mode == 0) decode_tile(P, tile, b, g, mt, nt);
    else decode_tile_plain(P, tile, b, mt, nt);
    const int p0 = (mt / P.nqt) * P.tp, q0 = (mt % P.nqt) * P.tq, r0 = nt * P.tn;
    double* wg = P.mode == 0 ? P.w + ((int64_t)(b * 3 + g)) * P.v * P.v * P.ldw
                             : P.w + (int64_t)(b / P.o_div) * P.out_s1 + (int64_t)(b % P.o_div) * P.out_s2;
    const int64_t pitch = P.mode == 0 ? (int64_t)P.ldw : (int64_t)P.ldw64;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int m = 16 * warp + 8 * mi + sg;
      int64_t pq;          // flattened p*v + q of this row (mode 1: the row index)
      bool row_ok;
      if (P.mode != 0) {
        pq = (int64_t)mt * kBM + m;
        row_ok = pq < (int64_t)P.v;
      } else if (P.flat) {
        pq = (int64_t)mt * kBM + m;
        row_ok = pq < (int64_t)P.v * P.v;
      } else {
        const int p = p0 + m / P.tq, qq = q0 + m % P.tq;
        pq = (int64_t)p * P.v + qq;
        row_ok = (m < P.rows_valid) && (p < P.v) && (qq < P.v);
      }
      double* row = wg + pq * pitch + r0;
      if (row_ok) {
#pragma unroll
        for (int ni = 0; ni < NFRAG - SKIP; ++ni) {
          // C fragment columns 2*kq, 2*kq+1 of the MMA -> tile columns 8*ni + sigma(2kq), sigma(2kq+1)
          const int c0 = 8 * ni + kq, c1 = 8 * ni + kq + 4;
          if (r0 + c0 < P.ncols) row[c0] = acc[mi][ni][0];
          if (r0 + c1 < P.ncols) row[c1] = acc[mi][ni][1];
        }
      }
    }
  }
}

template <int NFRAG>
__global__ void __launch_bounds__(kGemmThreads, 1)
w_contract_dmma_kernel(const __grid_constant__ CUtensorMap tmA_n,   // patch: A box (16, tq, tp, 1) | flat: A  box (16, 128, 1)
                       const __grid_constant__ CUtensorMap tmA_t,   // patch: A box (16, tp, tq, 1) | flat: AT box (16, 128, 1)
                       const __grid_constant__ CUtensorMap tmB,     // box (16, tn, 1)
                       const GemmParams P) {
  static_assert(NFRAG >= 1 && NFRAG <= kMaxNFrag, "NFRAG out of range");
  static_assert(Chunking<NFRAG>::kNum <= 6, "chunk dispatch covers at most 6 chunks");
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
    if (warp == kConsumerWarps && lane == 0) producer_loop(P, &tmA_n, &tmA_t, &tmB, smem, full_bar, empty_bar, a_bytes, b_bytes);
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
  const uint32_t b_lo_off = (uint32_t)kAStageBytes + (uint32_t)sg * 128u + (uint32_t)((kq ^ sg) << 4);
  // rows of this thread: m = 16*warp + 8*mi + sigma(g8); offsets inside the normal / transposed box
  uint32_t a_off_n[2], a_off_t[2];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi) {
    int m = 16 * warp + 8 * mi + sg;
    a_off_n[mi] = (uint32_t)m * 128u + (uint32_t)((kq ^ (m & 7)) << 4);
    int mm = m < P.rows_valid ? m : 0;
    int srow = P.flat ? m : (mm % P.tq) * P.tp + (mm / P.tq);   // row of (p,q) inside the term-1 box
    a_off_t[mi] = (uint32_t)srow * 128u + (uint32_t)((kq ^ (srow & 7)) << 4);
  }
  int stage = 0, sub = 0;
  uint32_t phase = 0;
  ConsumerRegs<NFRAG> R;
  if ((int)blockIdx.x < P.total_tiles) {
    // prologue: fragments of the very first k-block
    mbar_wait(&full_bar[0], 0);
    R.alo[0] = lds128(smem_base + a_off_n[0]);
    R.alo[1] = lds128(smem_base + a_off_n[1]);
    R.ahi[0] = lds128(smem_base + (a_off_n[0] ^ 64u));
    R.ahi[1] = lds128(smem_base + (a_off_n[1] ^ 64u));
    load_b_chunk<NFRAG, 0>(R.bb[0], smem_base + b_lo_off);
  }
  int tile = blockIdx.x;
  consume_tiles<NFRAG, 0>(P, tile, P.main_tiles, R, stage, phase, sub, full_bar, empty_bar, smem_base, b_lo_off, a_off_n,
                          a_off_t, warp, lane, sg, kq);
  consume_tiles<NFRAG, (NFRAG >= 2 ? 1 : 0)>(P, tile, P.total_tiles, R, stage, phase, sub, full_bar, empty_bar, smem_base,
                                            b_lo_off, a_off_n, a_off_t, warp, lane, sg, kq);
}

// host-side dispatch on the column-fragment count
typedef void (*GemmKernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmParams);

inline GemmKernelFn gemm_kernel_for(int nfrag) {
  switch (nfrag) {
    case 1: return w_contract_dmma_kernel<1>;
    case 2: return w_contract_dmma_kernel<2>;
    case 3: return w_contract_dmma_kernel<3>;
    case 4: return w_contract_dmma_kernel<4>;
    case 5: return w_contract_dmma_kernel<5>;
    case 6: return w_contract_dmma_kernel<6>;
    case 7: return w_contract_dmma_kernel<7>;
    case 8: return w_contract_dmma_kernel<8>;
    case 9: return w_contract_dmma_kernel<9>;
    case 10: return w_contract_dmma_kernel<10>;
    case 11: return w_contract_dmma_kernel<11>;
    case 12: return w_contract_dmma_kernel<12>;
    case 13: return w_contract_dmma_kernel<13>;
    case 14: return w_contract_dmma_kernel<14>;
    case 15: return w_contract_dmma_kernel<15>;
    case 16: return w_contract_dmma_kernel<16>;
    default: return nullptr;
  }
}

}  // namespace mpqc_t
