// The handle: one device's resident, re-laid-out (T) problem -- operand pool, tiling plan, tensor maps, work buffers,
// staging arena -- and the launch helpers of the two hot kernels.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "host_util.cuh"
#include "t_energy.cuh"
#include "w_contract.cuh"

// -------------------------------------------------------------------------------------------------
// handle
// -------------------------------------------------------------------------------------------------
struct mpqc_t_handle {
  int device = 0;
  int64_t o = 0, v = 0, Kp = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  // resident operands
  double *A = nullptr, *AT = nullptr, *B = nullptr, *GV = nullptr, *T1T = nullptr, *eps_occ = nullptr, *eps_vir = nullptr;
  uint8_t* tile_sets = nullptr;
  bool uploaded = false;
  // plan
  int tp = 0, tq = 0, tn = 0, nfrag = 0, npt = 0, nqt = 0, nnt = 0, ldw = 0, kblocks = 0;
  int flat = 0, nmt = 0, skip_last = 0;
  int ntile = 0, ntt = 0;
  CUtensorMap tmA_n, tmA_t, tmB;
  // work buffers
  int batch_cap = 0;
  double *W = nullptr, *partial = nullptr;
  int64_t units_cap = 0;
  int* triples_dev = nullptr;
  double* unit_e_dev = nullptr;
  // operand pool.  Resident mode: npanel == o, panel x lives in slot x.  Panel-cache mode (density-fitted inputs
  // whose A does not fit): npanel < o slots, panels A_x are built on demand from the three-centre factors by the
  // plain-GEMM mode of the W-contraction kernel and kept under LRU while the units are walked occupied-block-wise.
  int npanel = 0;
  bool panel_mode = false;
  int df_block = 0;                // requested occupied block edge of the panel walk (0: automatic)
  std::vector<int> slot_of;        // [o]  x -> slot, -1 when not resident
  std::vector<int> x_of_slot;      // [npanel]
  std::vector<int64_t> slot_stamp; // [npanel] last use (LRU)
  int64_t stamp = 0;
  int* slot_map_dev = nullptr;     // [o] device copy of slot_of, read by the kernel in panel mode
  double *XaiT = nullptr, *XabT = nullptr, *T2raw = nullptr;   // [o][v][Kx], [v][v][Kx], t2[v][v][o][o] (panel mode)
  int64_t Kx = 0;                  // padded auxiliary dimension roundup8(naux) (>= 16)
  int64_t panels_built = 0;
  // staging arena of the uploads (raw input copies, <ia|bc> slabs): one allocation that lives with the handle, bump
  // allocated per upload -- cudaMalloc/cudaFree of GBs per call were measured at tens of ms per GB on some hosts
  double* arena = nullptr;
  size_t arena_cap = 0, arena_used = 0;
};

namespace {

void free_work(mpqc_t_handle* h) {
  cudaFree(h->W);
  cudaFree(h->partial);
  h->W = h->partial = nullptr;
  h->batch_cap = 0;
}

int plan(mpqc_t_handle* h) {
  const int v = (int)h->v;
  // row patch (tp x tq) of a 128-row tile: maximise useful rows, prefer odd tp (bank-conflict-free
  // fragment reads of the transposed box, see w_contract.cuh)
  double best = -1.0;
  for (int tp = 1; tp <= std::min(v, kBM); ++tp) {
    int tq = std::min(v, kBM / tp);
    if (tq < 1) continue;
    tq = std::min(tq, 256);
    double tiles = std::ceil((double)v / tp) * std::ceil((double)v / tq);
    double eff = (double)v * v / (tiles * kBM);
    double score = eff * ((tp & 1) ? 1.0 : 0.97);
    if (score > best + 1e-12) {
      best = score;
      h->tp = tp;
      h->tq = tq;
    }
  }
  h->npt = (v + h->tp - 1) / h->tp;
  h->nqt = (v + h->tq - 1) / h->tq;
  h->nmt = h->flat ? (int)(((int64_t)v * v + kBM - 1) / kBM) : h->npt * h->nqt;
  // column tiles: F = ceil(v/8) fragments over nnt tiles of NFRAG fragments; the last tile may drop one
  const int F = (v + 7) / 8;
  h->nnt = (F + kMaxNFrag - 1) / kMaxNFrag;
  h->nfrag = (F + h->nnt - 1) / h->nnt;
  h->tn = h->nfrag * 8;
  h->skip_last = (h->nfrag >= 2 && h->nnt * h->nfrag - 1 >= F) ? 1 : 0;
  h->ldw = (int)roundup(v, 16);
  h->kblocks = (int)((h->Kp + kBK - 1) / kBK);
  h->ntile = (v + kET - 1) / kET;
  h->ntt = h->ntile * (h->ntile + 1) * (h->ntile + 2) / 6;
  return MPQC_T_OK;
}

int make_maps(mpqc_t_handle* h) {
  const uint64_t v = (uint64_t)h->v, o = (uint64_t)h->o, Kp = (uint64_t)h->Kp, np = (uint64_t)h->npanel;
  if (h->flat) {
    uint64_t dims[3] = {Kp, v * v, np};
    uint64_t str[2] = {Kp * 8, v * v * Kp * 8};
    uint32_t box[3] = {(uint32_t)kBK, (uint32_t)kBM, 1};
    MPQC_T_TRY(encode_map(&h->tmA_n, h->A, 3, dims, str, box));
    MPQC_T_TRY(encode_map(&h->tmA_t, h->AT, 3, dims, str, box));
  } else {
    uint64_t dims[4] = {Kp, v, v, np};
    uint64_t str[3] = {Kp * 8, v * Kp * 8, v * v * Kp * 8};
    uint32_t box_n[4] = {(uint32_t)kBK, (uint32_t)h->tq, (uint32_t)h->tp, 1};
    uint32_t box_t[4] = {(uint32_t)kBK, (uint32_t)h->tp, (uint32_t)h->tq, 1};
    MPQC_T_TRY(encode_map(&h->tmA_n, h->A, 4, dims, str, box_n));
    MPQC_T_TRY(encode_map(&h->tmA_t, h->A, 4, dims, str, box_t));
  }
  {
    uint64_t dims[3] = {Kp, v, o * o};
    uint64_t str[2] = {Kp * 8, v * Kp * 8};
    uint32_t box[3] = {(uint32_t)kBK, (uint32_t)h->tn, 1};
    MPQC_T_TRY(encode_map(&h->tmB, h->B, 3, dims, str, box));
  }
  return MPQC_T_OK;
}

int ensure_work(mpqc_t_handle* h, int batch) {
  if (batch <= h->batch_cap) return MPQC_T_OK;
  free_work(h);
  size_t wbytes = (size_t)batch * 3 * h->v * h->v * h->ldw * sizeof(double);
  MPQC_T_CUDA(cudaMalloc(&h->W, wbytes));
  MPQC_T_CUDA(cudaMalloc(&h->partial, (size_t)batch * h->ntt * sizeof(double)));
  h->batch_cap = batch;
  return MPQC_T_OK;
}

int ensure_units(mpqc_t_handle* h, int64_t n) {
  if (n <= h->units_cap) return MPQC_T_OK;
  cudaFree(h->triples_dev);
  cudaFree(h->unit_e_dev);
  h->triples_dev = nullptr;
  h->unit_e_dev = nullptr;
  h->units_cap = 0;
  MPQC_T_CUDA(cudaMalloc(&h->triples_dev, (size_t)n * 3 * sizeof(int)));
  MPQC_T_CUDA(cudaMalloc(&h->unit_e_dev, (size_t)n * sizeof(double)));
  h->units_cap = n;
  return MPQC_T_OK;
}

int auto_batch(const mpqc_t_handle* h) {
  if (const char* env = getenv("MPQC_T_BATCH")) {   // A/B runs only (scripts/sweep.py, bench.py)
    const int b = atoi(env);
    if (b > 0) return b;
  }
  int64_t tiles_per_triple = 3LL * h->nmt * h->nnt;
  // >= 256 waves of tiles per launch keeps the persistent grid's tail (half a tile per SM) and the per-launch gaps
  // small.  Measured (round 2): larger batches are monotonically better at every shape -- benzene 18.9 / 22.6 / 24.4 /
  // 26.0 TFLOP/s at batch 2 / 5 / 16 / 47 (scripts/sweep.py, profiles/r02_batch_sweep.txt); trimer bench on one box
  // 32.47 / 32.59 / 32.65 at batch 2 / 4 / 8 (profiles/r02_ab_batch.txt).  Small batches that would keep W in L2 for
  // the energy kernel lose more to launch gaps and tail waves than they gain.
  int64_t nb = (256LL * h->num_sms + tiles_per_triple - 1) / tiles_per_triple;
  nb = std::max<int64_t>(1, std::min<int64_t>(nb, 1024));
  // bound the W workspace to ~6 GB
  size_t per = (size_t)3 * h->v * h->v * h->ldw * sizeof(double);
  int64_t cap = std::max<int64_t>(1, (int64_t)((6ull << 30) / per));
  return (int)std::min(nb, cap);
}

// Unit enumeration (include/mpqc_t.h): i-major list of i >= j >= k without i == j == k.  Units are decoded by
// arithmetic -- nothing of size O(o^3) is ever materialised on the host.
struct UnitIndex {
  std::vector<int64_t> start;   // start[i] = first unit whose leading index is i; start[o] = number of units
  explicit UnitIndex(int64_t o) : start((size_t)o + 1) {
    int64_t u = 0;
    for (int64_t i = 0; i < o; ++i) {
      start[(size_t)i] = u;
      u += (i + 1) * (i + 2) / 2 - 1;   // (j,k) pairs with k <= j <= i, minus (i,i,i)
    }
    start[(size_t)o] = u;
  }
  int64_t count() const { return start.back(); }
  void triple(int64_t unit, int& i, int& j, int& k) const {
    const int64_t ii = (std::upper_bound(start.begin(), start.end(), unit) - start.begin()) - 1;
    const int64_t r = unit - start[(size_t)ii];          // position inside the i group: j(j+1)/2 + k
    int64_t jj = (int64_t)((std::sqrt(8.0 * (double)r + 1.0) - 1.0) * 0.5);
    while (jj * (jj + 1) / 2 > r) --jj;
    while ((jj + 1) * (jj + 2) / 2 <= r) ++jj;
    i = (int)ii;
    j = (int)jj;
    k = (int)(r - jj * (jj + 1) / 2);
  }
};

GemmParams gemm_params(const mpqc_t_handle* h, int nbatch, const int* triples_dev) {
  GemmParams P;
  P.v = (int)h->v;
  P.o = (int)h->o;
  P.Kp = (int)h->Kp;
  P.kblocks = h->kblocks;
  P.tp = h->tp;
  P.tq = h->tq;
  P.tn = h->tn;
  P.nfrag = h->nfrag;
  P.npt = h->npt;
  P.nqt = h->nqt;
  P.nnt = h->nnt;
  P.flat = h->flat;
  P.skip_last = h->skip_last;
  P.nmt = h->nmt;
  P.tiles_per_group = h->nmt * h->nnt;
  P.total_tiles = nbatch * 3 * P.tiles_per_group;
  P.main_tiles = nbatch * 3 * h->nmt * (h->nnt - h->skip_last);
  P.ldw = h->ldw;
  P.rows_valid = h->flat ? kBM : h->tp * h->tq;
  P.triples = triples_dev;
  P.w = h->W;
  P.a_slot = h->panel_mode ? h->slot_map_dev : nullptr;
  P.mode = 0;
  P.ncols = (int)h->v;
  P.l_div = P.l_mod = P.r_div = P.r_mod = P.o_div = 1;
  P.out_s1 = P.out_s2 = P.ldw64 = 0;
  return P;
}

int launch_gemm(mpqc_t_handle* h, int nbatch, const int* triples_dev) {
  GemmParams P = gemm_params(h, nbatch, triples_dev);
  int grid = std::min(h->num_sms, P.total_tiles);
  GemmKernelFn fn = gemm_kernel_for(h->nfrag);
  MPQC_T_CHECK(fn != nullptr, MPQC_T_ERR_INTERNAL, "no W-contraction kernel for this column-fragment count");
  fn<<<grid, kGemmThreads, kGemmSmemBytes, h->stream>>>(h->tmA_n, h->tmA_t, h->tmB, P);
  MPQC_T_CUDA(cudaGetLastError());
  return MPQC_T_OK;
}

int launch_energy(mpqc_t_handle* h, int nbatch, const int* triples_dev, double* unit_e_dev) {
  EnergyParams E;
  E.v = (int)h->v;
  E.o = (int)h->o;
  E.ldw = h->ldw;
  E.ntile = h->ntile;
  E.ntt = h->ntt;
  E.triples = triples_dev;
  E.w = h->W;
  E.gv = h->GV;
  E.t1t = h->T1T;
  E.eps_occ = h->eps_occ;
  E.eps_vir = h->eps_vir;
  E.tile_sets = h->tile_sets;
  E.partial = h->partial;
  t_energy_fused_kernel<<<dim3((unsigned)h->ntt, (unsigned)nbatch), kEThreads, kEnergySmemBytes, h->stream>>>(E);
  MPQC_T_CUDA(cudaGetLastError());
  t_energy_finish_kernel<<<nbatch, 256, 0, h->stream>>>(h->partial, h->ntt, triples_dev, unit_e_dev);
  MPQC_T_CUDA(cudaGetLastError());
  return MPQC_T_OK;
}

// staging arena: reserve once per upload (grows only), then bump-allocate 256-byte aligned pieces
int arena_reserve(mpqc_t_handle* h, size_t doubles) {
  h->arena_used = 0;
  if (doubles <= h->arena_cap) return MPQC_T_OK;
  cudaFree(h->arena);
  h->arena = nullptr;
  h->arena_cap = 0;
  MPQC_T_CUDA(cudaMalloc(&h->arena, std::max<size_t>(doubles, 32) * sizeof(double)));
  h->arena_cap = doubles;
  return MPQC_T_OK;
}

inline size_t arena_round(size_t doubles) { return (doubles + 31) / 32 * 32; }

double* arena_take(mpqc_t_handle* h, size_t doubles) {
  const size_t need = arena_round(doubles);
  if (h->arena_used + need > h->arena_cap) return nullptr;
  double* p = h->arena + h->arena_used;
  h->arena_used += need;
  return p;
}

// Allocates the operand pool A (and AT in flat mode) with `npanel` panel slots, decides the row mode, plans the tiling
// and encodes the tensor maps.  npanel == o: every panel resident (slot = x).  `extra_bytes`: what the caller will
// additionally keep on the device (staged factors ...), for the feasibility check.
int alloc_operands(mpqc_t_handle* h, int npanel, double extra_bytes) {
  const int64_t o = h->o, v = h->v;
  size_t free_b = 0, total_b = 0;
  MPQC_T_CUDA(cudaMemGetInfo(&free_b, &total_b));
  // a re-used handle (communicator cache) already holds a pool: count it as available, keep it if it still fits the plan
  const double held = (h->A ? 1.0 : 0.0) * (double)h->npanel * v * v * h->Kp * 8.0 * (h->AT ? 2.0 : 1.0);
  free_b += (size_t)held;
  // "flat" mode keeps a transposed copy AT of the big operand so both GEMM terms read 128 consecutive
  // flattened (p,q) rows (no row-patch padding).  Use it when 2|A| + the rest leaves >= 25% of free HBM.
  const double a_bytes = (double)npanel * v * v * h->Kp * 8.0;
  const double w_one = 3.0 * (double)v * v * (double)roundup(v, 16) * 8.0;
  const char* env = getenv("MPQC_T_FLAT");
  h->flat = (2.0 * a_bytes + extra_bytes + 8e9) < 0.75 * (double)free_b ? 1 : 0;
  if (env) h->flat = atoi(env) != 0;
  // feasibility: the accepted (o, v) range is far wider than what one device can hold.  Refuse here, with the numbers,
  // instead of failing inside some later cudaMalloc: operand pool + the W workspace of ONE triple + the caller's extras.
  const double need = a_bytes * (h->flat ? 2.0 : 1.0) + w_one + extra_bytes;
  if (need > 0.98 * (double)free_b) {
    char buf[360];
    snprintf(buf, sizeof(buf),
             "problem o=%lld v=%lld needs %.1f GB more device memory (operand panels %.1f GB in %d slots, W workspace "
             "%.1f GB per triple, staging %.1f GB) but only %.1f GB are free on device %d",
             (long long)o, (long long)v, need * 1e-9, a_bytes * (h->flat ? 2.0 : 1.0) * 1e-9, npanel, w_one * 1e-9,
             extra_bytes * 1e-9, (double)free_b * 1e-9, h->device);
    return fail(MPQC_T_ERR_OOM, buf, __FILE__, __LINE__);
  }
  const bool keep = h->A != nullptr && h->npanel == npanel && ((h->AT != nullptr) == (h->flat != 0));
  if (!keep) {
    cudaFree(h->A);
    cudaFree(h->AT);
    h->A = h->AT = nullptr;
    free_work(h);
    h->npanel = npanel;
    MPQC_T_CUDA(cudaMalloc(&h->A, (size_t)npanel * v * v * h->Kp * sizeof(double)));
    if (h->flat) MPQC_T_CUDA(cudaMalloc(&h->AT, (size_t)npanel * v * v * h->Kp * sizeof(double)));
  }
  h->panel_mode = npanel < o;
  h->panels_built = 0;
  cudaFree(h->slot_map_dev);
  h->slot_map_dev = nullptr;
  MPQC_T_TRY(plan(h));
  MPQC_T_TRY(make_maps(h));
  h->slot_of.assign((size_t)o, -1);
  h->x_of_slot.assign((size_t)npanel, -1);
  h->slot_stamp.assign((size_t)npanel, 0);
  if (!h->panel_mode)
    for (int64_t x = 0; x < o; ++x) h->slot_of[(size_t)x] = h->x_of_slot[(size_t)x] = (int)x;
  return MPQC_T_OK;
}

}  // namespace
