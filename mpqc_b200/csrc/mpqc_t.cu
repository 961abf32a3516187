// libmpqc_t_cuda.so -- host driver + C ABI (include/mpqc_t.h) of the B200 (T) path.
//
// Replaces the body of CCSD_T::compute_ccsd_t() (ccsd_t.h:144-177) / compute_ccsd_t_coarse_grain
// (ccsd_t.h:200-711): integrals and amplitudes arrive as dense buffers, are re-laid-out once into
// occupied-major operand panels (relayout.cuh), and the (i>=j>=k) triple space is walked in batches
// of {W-contraction DMMA kernel (w_contract.cuh) -> fused energy kernel (t_energy.cuh)}.
// There is no CPU fallback: without a CUDA device every entry point returns MPQC_T_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <dlfcn.h>

#include <functional>

#include "common.cuh"
#include "comm.cuh"
#include "microbench.cuh"
#include "relayout.cuh"
#include "t_energy.cuh"
#include "w_contract.cuh"

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


}  // namespace

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
  int64_t tiles_per_triple = 3LL * h->nmt * h->nnt;
  // >= 128 waves of tiles per launch keeps the persistent grid's tail (half a tile per SM) and the per-launch gaps
  // under ~0.5 %.  Measured (scripts/sweep.py, round 2): larger batches are monotonically better at every shape
  // (benzene 18.9 / 22.6 / 24.4 / 26.0 TFLOP/s at batch 2 / 5 / 16 / 47; trimer 32.36 / 32.55 / 32.64 at 1 / 2 / 4) --
  // small batches that would keep W in L2 for the energy kernel lose more to launch gaps and tail waves than they gain.
  int64_t nb = (128LL * h->num_sms + tiles_per_triple - 1) / tiles_per_triple;
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

// this worker's place in the (T) communicator; nranks == 1 means "no exchange"
struct CommView {
  int rank = 0, nranks = 1;
  NcclApi::comm_t comm = nullptr;
  double* scratch = nullptr;
};

inline size_t share_of(size_t n, int nranks) { return (n + (size_t)nranks - 1) / (size_t)nranks; }
inline size_t padded_count(size_t n, int nranks) { return share_of(n, nranks) * (size_t)nranks; }

// Puts host tensor src[n] into dst on EVERY rank (dst capacity >= padded_count(n, nranks)): this rank moves only its
// 1/nranks share over its own PCIe link, then one in-place ncclAllGather over NVLink completes the tensor.  With
// nranks == 1 it is a plain host->device copy.  The collective is issued even if the local copy failed, so that peer
// ranks never wait for a rank that dropped out.
int replicate_from_host(const CommView& cv, double* dst, const double* src, size_t n, cudaStream_t st, int64_t* h2d) {
  if (n == 0) return MPQC_T_OK;
  if (cv.nranks <= 1) {
    MPQC_T_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (h2d) *h2d += (int64_t)(n * sizeof(double));
    return MPQC_T_OK;
  }
  const size_t cnt = share_of(n, cv.nranks);
  const size_t lo = std::min(n, cnt * (size_t)cv.rank), hi = std::min(n, lo + cnt);
  int rc = MPQC_T_OK;
  if (hi > lo) {
    rc = cuda_status(cudaMemcpyAsync(dst + lo, src + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, st),
                     "cudaMemcpyAsync(host share)", __FILE__, __LINE__);
    if (h2d) *h2d += (int64_t)((hi - lo) * sizeof(double));
  }
  const NcclApi& nc = nccl_api();
  int r = nc.AllGather(dst + cnt * (size_t)cv.rank, dst, cnt, kNcclFloat64, cv.comm, st);
  if (rc == MPQC_T_OK) rc = nccl_status(r, "ncclAllGather(input replication)", __FILE__, __LINE__);
  return rc;
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

// host tensor -> device copy in the handle's staging arena (sharded + all-gathered when a communicator is present), or
// an alias of a device pointer
struct Staged {
  const double* ptr = nullptr;
};

inline size_t staged_size(size_t n, bool on_device, int nranks) { return on_device ? 0 : arena_round(padded_count(n, nranks)); }

int stage_in(mpqc_t_handle* h, Staged& s, const double* src, size_t n, bool on_device, const CommView& cv, cudaStream_t st,
             int64_t* h2d) {
  if (on_device) {
    s.ptr = src;
    return MPQC_T_OK;
  }
  double* dst = arena_take(h, padded_count(n, cv.nranks));
  MPQC_T_CHECK(dst != nullptr, MPQC_T_ERR_INTERNAL, "staging arena too small");
  MPQC_T_TRY(replicate_from_host(cv, dst, src, n, st, h2d));
  s.ptr = dst;
  return MPQC_T_OK;
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

int upload_impl(mpqc_t_handle* h, const mpqc_t_problem* p, bool on_device, const CommView& cv, mpqc_t_stats* stats) {
  const int64_t o = h->o, v = h->v, Kp = h->Kp;
  cudaStream_t st = h->stream;
  int64_t launches = 0, h2d = 0;
  const double t0 = now_s();
  double t_copy = 0.0;
  h->uploaded = false;
  // <ia|bc> streams through slabs of whole kap rows (host inputs)
  const size_t row = (size_t)v * v * o;  // doubles per kap
  const size_t slab_bytes = cv.nranks > 1 ? (size_t(2) << 30) : (size_t(1) << 30);
  const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(v, (int64_t)(slab_bytes / (row * 8 + 1)) + 1));
  // staging arena: raw copies of t1, t2, g_abij, g_aijk and two slabs (nothing when the inputs are on the device)
  const size_t arena_need = on_device ? 0
                                      : staged_size((size_t)v * o, false, 1) + 2 * staged_size((size_t)v * v * o * o, false, cv.nranks) +
                                            staged_size((size_t)v * o * o * o, false, cv.nranks) +
                                            2 * arena_round(padded_count((size_t)slab * row, cv.nranks));
  const double arena_new = arena_need > h->arena_cap ? (double)arena_need * 8.0 : 0.0;
  // dense inputs: all o panels resident
  MPQC_T_TRY(alloc_operands(h, (int)o, arena_new));
  MPQC_T_TRY(arena_reserve(h, arena_need));
  // Ordering contract (include/mpqc_t.h): device-resident inputs may have been produced on any stream of the caller;
  // the handle's stream is non-blocking, so wait for the whole device before reading them.
  if (on_device) MPQC_T_CUDA(cudaDeviceSynchronize());

  MPQC_T_CUDA(cudaMemsetAsync(h->A, 0, (size_t)o * v * v * Kp * sizeof(double), st));
  if (h->flat) MPQC_T_CUDA(cudaMemsetAsync(h->AT, 0, (size_t)o * v * v * Kp * sizeof(double), st));
  MPQC_T_CUDA(cudaMemsetAsync(h->B, 0, (size_t)o * o * v * Kp * sizeof(double), st));

  {
    double tc = now_s();
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    MPQC_T_CUDA(cudaMemcpyAsync(h->eps_occ, p->eps_occ, o * sizeof(double), kind, st));
    MPQC_T_CUDA(cudaMemcpyAsync(h->eps_vir, p->eps_vir, v * sizeof(double), kind, st));
    if (!on_device) h2d += (o + v) * 8;
    Staged t1, t2, gabij, gaijk;
    CommView solo;   // the tiny t1 is copied whole by every rank
    MPQC_T_TRY(stage_in(h, t1, p->t1, (size_t)v * o, on_device, solo, st, &h2d));
    MPQC_T_TRY(stage_in(h, t2, p->t2, (size_t)v * v * o * o, on_device, cv, st, &h2d));
    MPQC_T_TRY(stage_in(h, gabij, p->g_abij, (size_t)v * v * o * o, on_device, cv, st, &h2d));
    MPQC_T_TRY(stage_in(h, gaijk, p->g_aijk, (size_t)v * o * o * o, on_device, cv, st, &h2d));
    if (!on_device) {
      MPQC_T_CUDA(cudaStreamSynchronize(st));
      t_copy += now_s() - tc;
    }
    // T1T[i][a] = t1[a][i]
    MPQC_T_TRY(launch_transpose(st, t1.ptr, h->T1T, v, 1, o, 1, v, 0, 0, &launches));
    // GV[(i,j)][(a,b)] = g_abij[(a,b)][(i,j)]
    MPQC_T_TRY(launch_transpose(st, gabij.ptr, h->GV, v * v, 1, o * o, 1, v * v, 0, 0, &launches));
    // B particle part: t2[kap][r][(y,z)] -> B[(y,z)][r][kap]
    MPQC_T_TRY(launch_transpose(st, t2.ptr, h->B, v, v, o * o, 1, v * Kp, 0, Kp, &launches));
    // B hole part: g_aijk[r][(y,z)][l] -> B[(y,z)][r][v + l]
    MPQC_T_TRY(launch_copy_hole(st, gaijk.ptr, h->B, v, o * o, o, 1, Kp, 0, v * Kp, v, 1.0, &launches));
    // A hole part: -t2[(p,q)][x][l] -> A[x][p][q][v + l]   (and AT[x][q][p][v + l])
    MPQC_T_TRY(launch_copy_hole(st, t2.ptr, h->A, v * v, o, o, v, v * Kp, Kp, v * v * Kp, v, -1.0, &launches));
    if (h->flat)
      MPQC_T_TRY(launch_copy_hole(st, t2.ptr, h->AT, v * v, o, o, v, Kp, v * Kp, v * v * Kp, v, -1.0, &launches));
    MPQC_T_CUDA(cudaStreamSynchronize(st));
  }

  // A particle part: g_abci[kap][p][(q,x)] -> A[x][p][q][kap], streamed in kap slabs
  if (on_device) {
    MPQC_T_TRY(launch_transpose(st, p->g_abci, h->A, v, v, v * o, o, Kp, v * v * Kp, v * Kp, &launches));
    if (h->flat)   // AT[x][P][Q][d] = g_abci[d][Q][P][x]: mid (first virtual) -> Q, j / o (second virtual) -> P
      MPQC_T_TRY(launch_transpose(st, p->g_abci, h->AT, v, v, v * o, o, v * Kp, v * v * Kp, Kp, &launches));
    MPQC_T_CUDA(cudaStreamSynchronize(st));
  } else {
    // slabs of whole kap rows; each slab crosses PCIe once in total (1/nranks of it per rank) and is completed by an
    // all-gather, then transposed into place.  Two slabs so that the copy of the next one is queued while the
    // transposes of the current one run.
    struct { double* p; } buf[2];
    for (int s = 0; s < 2; ++s) {
      buf[s].p = arena_take(h, padded_count((size_t)slab * row, cv.nranks));
      MPQC_T_CHECK(buf[s].p != nullptr, MPQC_T_ERR_INTERNAL, "staging arena too small");
    }
    EventList events;               // copy (+ all-gather) time of every slab, device-timed on the handle's stream
    std::vector<cudaEvent_t> ev;
    int which = 0;
    for (int64_t d0 = 0; d0 < v; d0 += slab, which ^= 1) {
      const int64_t nd = std::min(slab, v - d0);
      cudaEvent_t e0, e1;
      MPQC_T_TRY(events.add(&e0));
      MPQC_T_TRY(events.add(&e1));
      MPQC_T_CUDA(cudaEventRecord(e0, st));
      MPQC_T_TRY(replicate_from_host(cv, buf[which].p, p->g_abci + (size_t)d0 * row, (size_t)nd * row, st, &h2d));
      MPQC_T_CUDA(cudaEventRecord(e1, st));
      ev.push_back(e0);
      ev.push_back(e1);
      MPQC_T_TRY(launch_transpose(st, buf[which].p, h->A + d0, nd, v, v * o, o, Kp, v * v * Kp, v * Kp, &launches));
      if (h->flat)
        MPQC_T_TRY(launch_transpose(st, buf[which].p, h->AT + d0, nd, v, v * o, o, v * Kp, v * v * Kp, Kp, &launches));
    }
    MPQC_T_CUDA(cudaStreamSynchronize(st));
    for (size_t q = 0; q + 1 < ev.size(); q += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[q], ev[q + 1]);
      t_copy += ms * 1e-3;
    }
  }
  MPQC_T_CUDA(cudaGetLastError());
  h->uploaded = true;
  if (stats) {
    double tot = now_s() - t0;
    stats->seconds_upload += t_copy;
    stats->seconds_relayout += tot - t_copy;
    stats->kernel_launches += launches;
    stats->bytes_h2d += h2d;
  }
  return MPQC_T_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Density-fitted inputs: the three integral classes are assembled on the device, straight into
// the operand layouts, from the three-centre factors (what the reference's [df] formulas evaluate through TiledArray on
// the host, ccsd_t.h:2210-2244 with is_df()) -- by the SAME TMA + DMMA pipeline as the W contraction, in its plain
// batched NT-GEMM mode (w_contract.cuh, GemmParams::mode = 1).  No library GEMM is involved.
//
//   A[x][p][q][kap<v] = <x kap|p q> = sum_K Xai[K,p,x] Xab[K,kap,q]      C_(x,q)[p][kap],   L = XaiT[x], R = XabT[q]
//   AT[x][p][q][kap]  = A[x][q][p][kap]                                  C_(x,p)[q][kap],   L = XaiT[x], R = XabT[p]
//   GV[i][j][a][b]    = <ij|ab>     = sum_K Xai[K,a,i] Xai[K,b,j]        C_(i,j)[a][b],     L = XaiT[i], R = XaiT[j]
//   B[y][z][r][v+l]   = <yz|lr>     = sum_K Xai[K,r,z] Xij[K,y,l]        C_(y,z)[r][l],     L = XaiT[z], R = XijT[y]
// with the factor copies XaiT[x][a][K], XabT[q][kap][K] = Xab[K][kap][q], XijT[y][l][K] (K fastest, zero padded to Kx).
// ---------------------------------------------------------------------------------------------------------------
struct PlainGemm {
  const double* L;   // [l_batches][M][Kx]
  int64_t l_batches, M;
  const double* R;   // [r_batches][N][Kx]
  int64_t r_batches, N;
  int64_t Kx;
  int nbatch, l_div, l_mod, r_div, r_mod, o_div;
  double* out;
  int64_t out_s1, out_s2, ldw;
};

int launch_plain_gemm(mpqc_t_handle* h, const PlainGemm& g, int64_t* launches) {
  if (g.nbatch <= 0 || g.M <= 0 || g.N <= 0) return MPQC_T_OK;
  const int F = (int)((g.N + 7) / 8);
  const int nnt = (F + kMaxNFrag - 1) / kMaxNFrag;
  const int nfrag = (F + nnt - 1) / nnt;
  const int tn = nfrag * 8;
  const int skip_last = (nfrag >= 2 && nnt * nfrag - 1 >= F) ? 1 : 0;
  const int nmt = (int)((g.M + kBM - 1) / kBM);
  MPQC_T_CHECK((int64_t)g.nbatch * nmt * nnt < (1LL << 31), MPQC_T_ERR_INTERNAL, "too many tiles in one factor GEMM");
  CUtensorMap tmL, tmR;
  {
    uint64_t dims[3] = {(uint64_t)g.Kx, (uint64_t)g.M, (uint64_t)g.l_batches};
    uint64_t str[2] = {(uint64_t)g.Kx * 8, (uint64_t)g.M * g.Kx * 8};
    uint32_t box[3] = {(uint32_t)kBK, (uint32_t)kBM, 1};
    MPQC_T_TRY(encode_map(&tmL, const_cast<double*>(g.L), 3, dims, str, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)g.Kx, (uint64_t)g.N, (uint64_t)g.r_batches};
    uint64_t str[2] = {(uint64_t)g.Kx * 8, (uint64_t)g.N * g.Kx * 8};
    uint32_t box[3] = {(uint32_t)kBK, (uint32_t)tn, 1};
    MPQC_T_TRY(encode_map(&tmR, const_cast<double*>(g.R), 3, dims, str, box));
  }
  GemmParams P;
  memset(&P, 0, sizeof(P));
  P.v = (int)g.M;
  P.o = 0;
  P.Kp = (int)g.Kx;
  P.kblocks = (int)((g.Kx + kBK - 1) / kBK);
  P.tp = P.tq = 1;
  P.npt = P.nqt = 1;
  P.tn = tn;
  P.nfrag = nfrag;
  P.nnt = nnt;
  P.skip_last = skip_last;
  P.flat = 1;
  P.nmt = nmt;
  P.tiles_per_group = nmt * nnt;
  P.total_tiles = g.nbatch * nmt * nnt;
  P.main_tiles = g.nbatch * nmt * (nnt - skip_last);
  P.rows_valid = kBM;
  P.w = g.out;
  P.mode = 1;
  P.ncols = (int)g.N;
  P.l_div = g.l_div;
  P.l_mod = g.l_mod;
  P.r_div = g.r_div;
  P.r_mod = g.r_mod;
  P.o_div = g.o_div;
  P.out_s1 = g.out_s1;
  P.out_s2 = g.out_s2;
  P.ldw64 = g.ldw;
  GemmKernelFn fn = gemm_kernel_for(nfrag);
  MPQC_T_CHECK(fn != nullptr, MPQC_T_ERR_INTERNAL, "no GEMM kernel for this column-fragment count");
  MPQC_T_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
  const int grid = std::min(h->num_sms, P.total_tiles);
  fn<<<grid, kGemmThreads, kGemmSmemBytes, h->stream>>>(tmL, tmL, tmR, P);
  MPQC_T_CUDA(cudaGetLastError());
  if (launches) ++*launches;
  return MPQC_T_OK;
}

// Operand panels of occupied indices x0 .. x0+nx-1 into pool slots slot0 .. (consecutive).  The particle part is ONE
// plain GEMM per call whose rows are the flattened (x, p) pairs of all nx panels (no row padding per panel: consecutive
// slots are v rows of pitch v*Kp apart, so row m = (x - x0)*v + p lands at slot0*pstride + m*(v*Kp)), batched over q;
// in panel-cache mode the hole part is copied from t2; in flat mode the transposed copy AT_x is then a row-wise
// transposing copy of the finished panel (HBM-bound, ~10x cheaper than a second GEMM).  In resident mode the caller has
// written the hole part of A before.
int build_panels(mpqc_t_handle* h, int x0, int nx, int slot0, int64_t* launches) {
  const int64_t o = h->o, v = h->v, Kp = h->Kp;
  const int64_t pstride = v * v * Kp;
  PlainGemm g;
  g.L = h->XaiT + (int64_t)x0 * v * h->Kx;   // rows (x, p), x = x0 .. x0+nx-1
  g.l_batches = 1;
  g.M = (int64_t)nx * v;
  g.R = h->XabT;                             // batch entry q: R_q[kap][K] = Xab[K][kap][q]
  g.r_batches = v;
  g.N = v;
  g.Kx = h->Kx;
  g.nbatch = (int)v;
  g.l_div = 1;
  g.l_mod = 1;
  g.r_div = 1;
  g.r_mod = (int)v;
  g.o_div = 1;
  g.out = h->A + (int64_t)slot0 * pstride;   // C_q[(x,p)][kap] -> A[slot][p][q][kap]
  g.out_s1 = Kp;
  g.out_s2 = 0;
  g.ldw = v * Kp;
  MPQC_T_TRY(launch_plain_gemm(h, g, launches));
  const unsigned cblocks = (unsigned)std::min<int64_t>((v * v * o + 255) / 256, 148 * 32);
  const unsigned tblocks = (unsigned)std::min<int64_t>((v * v + 7) / 8, 148 * 16);
  for (int x = x0; x < x0 + nx; ++x) {
    const int s = slot0 + (x - x0);
    if (h->panel_mode) {
      copy_hole_panel_kernel<<<cblocks, 256, 0, h->stream>>>(h->T2raw, h->A + (int64_t)s * pstride, v, o, x, Kp, 0);
      MPQC_T_CUDA(cudaGetLastError());
      if (launches) ++*launches;
    }
    if (h->flat) {
      transpose_panel_kernel<<<tblocks, 256, 0, h->stream>>>(h->A + (int64_t)s * pstride, h->AT + (int64_t)s * pstride, v, Kp);
      MPQC_T_CUDA(cudaGetLastError());
      if (launches) ++*launches;
    }
  }
  h->panels_built += nx;
  return MPQC_T_OK;
}

// occupied block edge of the panel walk: the pool must hold the panels of three occupied blocks
int panel_block_edge(const mpqc_t_handle* h) { return std::max(1, h->npanel / 3); }

int upload_df_impl(mpqc_t_handle* h, const mpqc_t_df_problem* p, bool on_device, const CommView& cv, mpqc_t_stats* stats) {
  const int64_t o = h->o, v = h->v, Kp = h->Kp, naux = p->naux;
  cudaStream_t st = h->stream;
  int64_t launches = 0, h2d = 0;
  const double t0 = now_s();
  double t_copy = 0.0;
  h->uploaded = false;
  if (on_device) MPQC_T_CUDA(cudaDeviceSynchronize());   // ordering contract for device-resident inputs (mpqc_t.h)
  const int64_t Kx = std::max<int64_t>(16, roundup(naux, 8));
  h->Kx = Kx;
  cudaFree(h->XaiT);
  cudaFree(h->XabT);
  cudaFree(h->T2raw);
  h->XaiT = h->XabT = h->T2raw = nullptr;

  // ---- resident or panel cache?  Resident when the whole operand fits beside everything else; otherwise the largest
  //      occupied block edge (<= 8) whose 3 blocks of panels fit.  MPQC_T_DF_BLOCK / mpqc_t_set_df_block force it. ----
  int block = h->df_block;
  if (const char* env = getenv("MPQC_T_DF_BLOCK")) block = atoi(env);
  const double panel_bytes = (double)v * v * Kp * 8.0;
  const double factors = ((double)o * v + (double)v * v) * Kx * 8.0;
  const double staging = on_device ? 0.0 : ((double)naux * v * v + (double)naux * v * o) * 8.0;   // raw factor copies
  const double t2_bytes = (double)v * v * o * o * 8.0;
  const double w_one = 3.0 * (double)v * v * (double)roundup(v, 16) * 8.0;
  size_t free_b = 0, total_b = 0;
  MPQC_T_CUDA(cudaMemGetInfo(&free_b, &total_b));
  free_b += (size_t)((h->A ? 1.0 : 0.0) * (double)h->npanel * panel_bytes * (h->AT ? 2.0 : 1.0));   // a re-used pool
  int npanel = (int)o;
  if (block > 0) {
    npanel = (int)std::min<int64_t>(o, 3LL * block);
  } else if (block == 0) {
    const double resident = (double)o * panel_bytes + factors + staging + (on_device ? 0.0 : t2_bytes) + w_one;
    if (resident > 0.80 * (double)free_b) {
      const double room = 0.80 * (double)free_b - (factors + staging + t2_bytes + w_one);
      const int fit = (int)std::floor(room / panel_bytes);
      npanel = (int)std::min<int64_t>(o, std::max(3, std::min(24, fit / 3 * 3)));
    }
  }
  const bool panel_mode = npanel < o;
  MPQC_T_TRY(alloc_operands(h, npanel, factors + staging + (panel_mode || !on_device ? t2_bytes : 0.0)));
  MPQC_T_CUDA(cudaMemsetAsync(h->A, 0, (size_t)npanel * v * v * Kp * sizeof(double), st));
  if (h->flat) MPQC_T_CUDA(cudaMemsetAsync(h->AT, 0, (size_t)npanel * v * v * Kp * sizeof(double), st));
  MPQC_T_CUDA(cudaMemsetAsync(h->B, 0, (size_t)o * o * v * Kp * sizeof(double), st));

  const double tc = now_s();
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  MPQC_T_CUDA(cudaMemcpyAsync(h->eps_occ, p->eps_occ, o * sizeof(double), kind, st));
  MPQC_T_CUDA(cudaMemcpyAsync(h->eps_vir, p->eps_vir, v * sizeof(double), kind, st));
  if (!on_device) h2d += (o + v) * 8;
  DevBuf xijt;
  {
    Staged t1, t2, xab, xij, xai;
    CommView solo;
    MPQC_T_TRY(arena_reserve(h, staged_size((size_t)v * o, on_device, 1) + staged_size((size_t)v * v * o * o, on_device, cv.nranks) +
                                    staged_size((size_t)naux * v * v, on_device, cv.nranks) +
                                    staged_size((size_t)naux * o * o, on_device, 1) +
                                    staged_size((size_t)naux * v * o, on_device, cv.nranks)));
    MPQC_T_TRY(stage_in(h, t1, p->t1, (size_t)v * o, on_device, solo, st, &h2d));
    MPQC_T_TRY(stage_in(h, t2, p->t2, (size_t)v * v * o * o, on_device, cv, st, &h2d));
    MPQC_T_TRY(stage_in(h, xab, p->x_ab, (size_t)naux * v * v, on_device, cv, st, &h2d));
    MPQC_T_TRY(stage_in(h, xij, p->x_ij, (size_t)naux * o * o, on_device, solo, st, &h2d));
    MPQC_T_TRY(stage_in(h, xai, p->x_ai, (size_t)naux * v * o, on_device, cv, st, &h2d));
    if (!on_device) {
      MPQC_T_CUDA(cudaStreamSynchronize(st));
      t_copy += now_s() - tc;
    }
    // factor copies with the auxiliary index fastest (one 128-byte TMA box row per 16 K), zero padded to Kx
    MPQC_T_CUDA(cudaMalloc(&h->XaiT, (size_t)o * v * Kx * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->XabT, (size_t)v * v * Kx * sizeof(double)));
    MPQC_T_TRY(xijt.alloc((size_t)o * o * Kx));
    MPQC_T_CUDA(cudaMemsetAsync(h->XaiT, 0, (size_t)o * v * Kx * sizeof(double), st));
    MPQC_T_CUDA(cudaMemsetAsync(h->XabT, 0, (size_t)v * v * Kx * sizeof(double), st));
    MPQC_T_CUDA(cudaMemsetAsync(xijt.p, 0, (size_t)o * o * Kx * sizeof(double), st));
    // XaiT[x][a][K] = Xai[K][a][x]:      in[kap=K][mid=a][j=x]    -> out[x * v*Kx + a * Kx + K]
    MPQC_T_TRY(launch_transpose(st, xai.ptr, h->XaiT, naux, v, o, 1, v * Kx, 0, Kx, &launches));
    // XabT[q][kap][K] = Xab[K][kap][q]:  in[kap=K][mid=kap][j=q]  -> out[q * v*Kx + kap * Kx + K]
    MPQC_T_TRY(launch_transpose(st, xab.ptr, h->XabT, naux, v, v, 1, v * Kx, 0, Kx, &launches));
    // XijT[y][l][K] = Xij[K][y][l]:      in[kap=K][mid=y][j=l]    -> out[y * o*Kx + l * Kx + K]
    MPQC_T_TRY(launch_transpose(st, xij.ptr, xijt.p, naux, o, o, 1, Kx, 0, o * Kx, &launches));

    // amplitude parts (same as the dense upload)
    MPQC_T_TRY(launch_transpose(st, t1.ptr, h->T1T, v, 1, o, 1, v, 0, 0, &launches));
    MPQC_T_TRY(launch_transpose(st, t2.ptr, h->B, v, v, o * o, 1, v * Kp, 0, Kp, &launches));
    if (panel_mode) {
      // the hole part of a panel is written when the panel is built: keep t2 on the device
      MPQC_T_CUDA(cudaMalloc(&h->T2raw, (size_t)v * v * o * o * sizeof(double)));
      MPQC_T_CUDA(cudaMemcpyAsync(h->T2raw, t2.ptr, (size_t)v * v * o * o * sizeof(double), cudaMemcpyDeviceToDevice, st));
      MPQC_T_CUDA(cudaMalloc(&h->slot_map_dev, (size_t)o * sizeof(int)));
    } else {   // resident: hole part of every panel now; AT is copied from the finished panels in build_panels
      MPQC_T_TRY(launch_copy_hole(st, t2.ptr, h->A, v * v, o, o, v, v * Kp, Kp, v * v * Kp, v, -1.0, &launches));
    }
    MPQC_T_CUDA(cudaStreamSynchronize(st));
  }

  PlainGemm g;
  // GV[i][j][a][b]
  g.L = h->XaiT; g.l_batches = o; g.M = v;
  g.R = h->XaiT; g.r_batches = o; g.N = v;
  g.Kx = Kx;
  g.nbatch = (int)(o * o);
  g.l_div = (int)o; g.l_mod = (int)o; g.r_div = 1; g.r_mod = (int)o; g.o_div = 1;
  g.out = h->GV; g.out_s1 = v * v; g.out_s2 = 0; g.ldw = v;
  MPQC_T_TRY(launch_plain_gemm(h, g, &launches));
  // B[y][z][r][v + l]
  g.L = h->XaiT; g.l_batches = o; g.M = v;
  g.R = xijt.p; g.r_batches = o; g.N = o;
  g.nbatch = (int)(o * o);
  g.l_div = 1; g.l_mod = (int)o; g.r_div = (int)o; g.r_mod = (int)o; g.o_div = 1;
  g.out = h->B + v; g.out_s1 = v * Kp; g.out_s2 = 0; g.ldw = Kp;
  MPQC_T_TRY(launch_plain_gemm(h, g, &launches));
  // operand panels: all of them now (resident), or on demand while the units are walked (panel cache)
  if (!panel_mode) {
    for (int x0 = 0; x0 < (int)o; x0 += 16) MPQC_T_TRY(build_panels(h, x0, std::min(16, (int)o - x0), x0, &launches));
  }
  MPQC_T_CUDA(cudaStreamSynchronize(st));
  MPQC_T_CUDA(cudaGetLastError());
  if (!panel_mode) {   // the factor copies are only needed again in panel mode
    cudaFree(h->XaiT);
    cudaFree(h->XabT);
    h->XaiT = h->XabT = nullptr;
  }
  h->uploaded = true;
  if (stats) {
    double tot = now_s() - t0;
    stats->seconds_upload += t_copy;
    stats->seconds_relayout += tot - t_copy;
    stats->kernel_launches += launches;
    stats->bytes_h2d += h2d;
  }
  return MPQC_T_OK;
}

// Panel-cache mode: make the panels of the occupied indices in `need` (sorted, unique) resident, evicting the least
// recently used panels that are not needed now, and refresh the device slot map.  All on the handle's stream, so the
// kernels that still read an evicted slot have finished before it is overwritten.
int ensure_panels(mpqc_t_handle* h, const std::vector<int>& need, int64_t* launches) {
  MPQC_T_CHECK((int)need.size() <= h->npanel, MPQC_T_ERR_INTERNAL, "panel pool smaller than one unit group");
  std::vector<char> wanted((size_t)h->o, 0);
  for (int x : need) wanted[(size_t)x] = 1;
  ++h->stamp;
  bool changed = false;
  for (int x : need) {
    int s = h->slot_of[(size_t)x];
    if (s < 0) {
      // victim: a free slot, else the least recently used slot whose panel is not needed by this group
      int victim = -1;
      for (int q = 0; q < h->npanel; ++q) {
        const int xq = h->x_of_slot[(size_t)q];
        if (xq < 0) { victim = q; break; }
        if (wanted[(size_t)xq]) continue;
        if (victim < 0 || h->slot_stamp[(size_t)q] < h->slot_stamp[(size_t)victim]) victim = q;
      }
      MPQC_T_CHECK(victim >= 0, MPQC_T_ERR_INTERNAL, "no evictable panel slot");
      if (h->x_of_slot[(size_t)victim] >= 0) h->slot_of[(size_t)h->x_of_slot[(size_t)victim]] = -1;
      h->x_of_slot[(size_t)victim] = x;
      h->slot_of[(size_t)x] = victim;
      MPQC_T_TRY(build_panels(h, x, 1, victim, launches));
      s = victim;
      changed = true;
    }
    h->slot_stamp[(size_t)s] = h->stamp;
  }
  if (changed)   // pageable source: the runtime stages it before returning, so slot_of may change again right away
    MPQC_T_CUDA(cudaMemcpyAsync(h->slot_map_dev, h->slot_of.data(), (size_t)h->o * sizeof(int), cudaMemcpyHostToDevice,
                                h->stream));
  return MPQC_T_OK;
}

// key of the occupied-block triple a unit belongs to (block edge bo): units with equal keys need at most 3 bo panels
inline int64_t block_key(int i, int j, int k, int bo) {
  const int64_t nb = 4096 / bo + 2;
  return ((int64_t)(i / bo) * nb + (j / bo)) * nb + (k / bo);
}

// Panel-cache mode of run_units: the units are processed grouped by occupied-block triple (sorted by key, so
// consecutive groups share their leading blocks and the LRU pool keeps those panels); before a group runs, the
// panels it needs are built on the stream.  Results return in the caller's unit order.
int run_units_panels(mpqc_t_handle* h, const std::vector<int>& tri, int64_t n, int batch, double* unit_e_host,
                     mpqc_t_stats* stats, double* vblock_dev) {
  const int bo = panel_block_edge(h);
  std::vector<int64_t> order((size_t)n);
  for (int64_t u = 0; u < n; ++u) order[(size_t)u] = u;
  std::vector<int64_t> key((size_t)n);
  for (int64_t u = 0; u < n; ++u) key[(size_t)u] = block_key(tri[3 * u], tri[3 * u + 1], tri[3 * u + 2], bo);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[(size_t)a] < key[(size_t)b]; });
  std::vector<int> tri_sorted((size_t)n * 3);
  for (int64_t q = 0; q < n; ++q)
    for (int c = 0; c < 3; ++c) tri_sorted[3 * q + c] = tri[3 * order[(size_t)q] + c];
  MPQC_T_CUDA(cudaMemcpyAsync(h->triples_dev, tri_sorted.data(), tri_sorted.size() * sizeof(int), cudaMemcpyHostToDevice,
                              h->stream));
  EventList events;
  cudaEvent_t e_begin, e_end;
  MPQC_T_TRY(events.add(&e_begin));
  MPQC_T_TRY(events.add(&e_end));
  MPQC_T_CUDA(cudaEventRecord(e_begin, h->stream));
  int64_t launches = 0;
  const int64_t built0 = h->panels_built;
  for (int64_t g0 = 0; g0 < n;) {
    int64_t g1 = g0;
    std::vector<int> need;
    while (g1 < n && key[(size_t)order[(size_t)g1]] == key[(size_t)order[(size_t)g0]]) {
      for (int c = 0; c < 3; ++c) need.push_back(tri_sorted[3 * g1 + c]);
      ++g1;
    }
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    MPQC_T_TRY(ensure_panels(h, need, &launches));
    for (int64_t off = g0; off < g1; off += batch) {
      const int nb = (int)std::min<int64_t>(batch, g1 - off);
      MPQC_T_TRY(launch_gemm(h, nb, h->triples_dev + 3 * off));
      MPQC_T_TRY(launch_energy(h, nb, h->triples_dev + 3 * off, h->unit_e_dev + off));
      launches += 3;
      if (vblock_dev) {
        t_energy_vblock_kernel<<<(h->ntt + 255) / 256, 256, 0, h->stream>>>(h->partial, h->ntt, nb,
                                                                          h->triples_dev + 3 * off, vblock_dev);
        MPQC_T_CUDA(cudaGetLastError());
        ++launches;
      }
    }
    g0 = g1;
  }
  MPQC_T_CUDA(cudaEventRecord(e_end, h->stream));
  std::vector<double> ue((size_t)n);
  MPQC_T_CUDA(cudaMemcpyAsync(ue.data(), h->unit_e_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  MPQC_T_CUDA(cudaStreamSynchronize(h->stream));
  MPQC_T_CUDA(cudaGetLastError());
  for (int64_t q = 0; q < n; ++q) unit_e_host[order[(size_t)q]] = ue[(size_t)q];
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e_begin, e_end);
    stats->seconds_compute += ms * 1e-3;
    stats->units += n;
    stats->kernel_launches += launches;
    stats->flops += (double)n * mpqc_t_unit_flops(h->o, h->v);
    const double mpad = (double)h->nmt * kBM, npad = (double)h->nnt * h->tn - 8.0 * h->skip_last;
    // executed: the triples themselves + the panels built for them (2 naux v^3 each)
    stats->flops_executed += (double)n * 3.0 * 2.0 * 2.0 * mpad * npad * (double)h->Kp +
                             (double)(h->panels_built - built0) * 2.0 * (double)h->Kx * (double)h->v * h->v * h->v;
    stats->bytes_d2h += n * 8;
    stats->bytes_h2d += n * 12;
  }
  return MPQC_T_OK;
}

// Run an explicit list of units (indices into the global enumeration).  unit_e_host[n] receives the
// weighted per-unit energies.  vblock_dev (optional, [ntt] on the device, zeroed by the caller) accumulates the
// decomposition of the same energy over virtual-block triples.  Synchronises the stream before returning.
int run_units(mpqc_t_handle* h, const UnitIndex& ux, const int64_t* units, int64_t n, int batch,
              double* unit_e_host, mpqc_t_stats* stats, bool profile, double* vblock_dev = nullptr) {
  if (n == 0) return MPQC_T_OK;
  MPQC_T_CUDA(cudaSetDevice(h->device));
  if (batch <= 0) batch = auto_batch(h);
  batch = (int)std::min<int64_t>(batch, n);
  batch = std::min(batch, 65535);
  {  // tile indices are 32-bit
    const int64_t tiles_per_triple = 3LL * h->nmt * h->nnt;
    batch = (int)std::max<int64_t>(1, std::min<int64_t>(batch, ((1LL << 31) - 1) / tiles_per_triple));
  }
  MPQC_T_TRY(ensure_work(h, batch));
  MPQC_T_TRY(ensure_units(h, n));
  std::vector<int> tri((size_t)n * 3);
  for (int64_t u = 0; u < n; ++u) ux.triple(units[u], tri[3 * u], tri[3 * u + 1], tri[3 * u + 2]);
  if (h->panel_mode) return run_units_panels(h, tri, n, batch, unit_e_host, stats, vblock_dev);
  MPQC_T_CUDA(cudaMemcpyAsync(h->triples_dev, tri.data(), tri.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));

  const int64_t nbatches = (n + batch - 1) / batch;
  profile = profile && nbatches <= 8192;
  EventList events;
  std::vector<cudaEvent_t> ev;
  cudaEvent_t e_begin, e_end;
  MPQC_T_TRY(events.add(&e_begin));
  MPQC_T_TRY(events.add(&e_end));
  if (profile) {
    ev.resize((size_t)nbatches * 3);
    for (auto& e : ev) MPQC_T_TRY(events.add(&e));
  }
  MPQC_T_CUDA(cudaEventRecord(e_begin, h->stream));
  int64_t launches = 0;
  for (int64_t bi = 0; bi < nbatches; ++bi) {
    const int64_t off = bi * batch;
    const int nb = (int)std::min<int64_t>(batch, n - off);
    if (profile) cudaEventRecord(ev[3 * bi], h->stream);
    MPQC_T_TRY(launch_gemm(h, nb, h->triples_dev + 3 * off));
    if (profile) cudaEventRecord(ev[3 * bi + 1], h->stream);
    MPQC_T_TRY(launch_energy(h, nb, h->triples_dev + 3 * off, h->unit_e_dev + off));
    if (profile) cudaEventRecord(ev[3 * bi + 2], h->stream);
    launches += 3;
    if (vblock_dev) {
      t_energy_vblock_kernel<<<(h->ntt + 255) / 256, 256, 0, h->stream>>>(h->partial, h->ntt, nb,
                                                                        h->triples_dev + 3 * off, vblock_dev);
      MPQC_T_CUDA(cudaGetLastError());
      ++launches;
    }
  }
  MPQC_T_CUDA(cudaEventRecord(e_end, h->stream));
  MPQC_T_CUDA(cudaMemcpyAsync(unit_e_host, h->unit_e_dev, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  MPQC_T_CUDA(cudaStreamSynchronize(h->stream));
  MPQC_T_CUDA(cudaGetLastError());
  if (stats) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e_begin, e_end);
    stats->seconds_compute += ms * 1e-3;
    if (profile) {
      double tg = 0, te = 0;
      for (int64_t bi = 0; bi < nbatches; ++bi) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, ev[3 * bi], ev[3 * bi + 1]);
        cudaEventElapsedTime(&b, ev[3 * bi + 1], ev[3 * bi + 2]);
        tg += a * 1e-3;
        te += b * 1e-3;
      }
      stats->seconds_contract += tg;
      stats->seconds_energy += te;
    }
    stats->units += n;
    stats->kernel_launches += launches;
    stats->flops += (double)n * mpqc_t_unit_flops(h->o, h->v);
    double mpad = (double)h->nmt * kBM, npad = (double)h->nnt * h->tn - 8.0 * h->skip_last;
    stats->flops_executed += (double)n * 3.0 * 2.0 * 2.0 * mpad * npad * (double)h->Kp;
    stats->bytes_d2h += n * 8;
    stats->bytes_h2d += n * 12;
  }
  return MPQC_T_OK;
}

// sum-all-reduce of n doubles held on the host through the member's pre-allocated device scratch (chunked), so the
// collective itself never allocates.  Result overwrites x on every rank.
int allreduce_host_vector(const CommView& cv, double* x, size_t n, cudaStream_t st) {
  const NcclApi& nc = nccl_api();
  for (size_t c0 = 0; c0 < n; c0 += kCommScratchDoubles) {
    const size_t cn = std::min(kCommScratchDoubles, n - c0);
    MPQC_T_CUDA(cudaMemcpyAsync(cv.scratch, x + c0, cn * sizeof(double), cudaMemcpyHostToDevice, st));
    MPQC_T_NCCL(nc.AllReduce(cv.scratch, cv.scratch, cn, kNcclFloat64, kNcclSum, cv.comm, st));
    MPQC_T_CUDA(cudaMemcpyAsync(x + c0, cv.scratch, cn * sizeof(double), cudaMemcpyDeviceToHost, st));
    MPQC_T_CUDA(cudaStreamSynchronize(st));
  }
  return MPQC_T_OK;
}

// Agreement on a status among all ranks: returns the number of ranks that reported a failure (or -1 when the
// collective itself failed).  Every rank calls it at the same points, whatever happened locally, so a rank that ran
// out of memory makes the others return an error instead of leaving them blocked in a later collective.
int count_failed_ranks(const CommView& cv, int local_rc, cudaStream_t st) {
  if (cv.nranks <= 1) return local_rc != MPQC_T_OK ? 1 : 0;
  double flag = local_rc != MPQC_T_OK ? 1.0 : 0.0;
  const std::string keep = last_error_string();
  int rc = allreduce_host_vector(cv, &flag, 1, st);
  if (local_rc != MPQC_T_OK) last_error_string() = keep;
  if (rc != MPQC_T_OK) return -1;
  return (int)(flag + 0.5);
}

int validate_problem(const mpqc_t_problem* p) {
  MPQC_T_CHECK(p != nullptr, MPQC_T_ERR_BAD_ARG, "problem is NULL");
  MPQC_T_CHECK(p->o >= 1 && p->v >= 1, MPQC_T_ERR_BAD_ARG, "o and v must be >= 1");
  MPQC_T_CHECK(p->o <= 4096 && p->v <= 2040, MPQC_T_ERR_BAD_ARG, "o <= 4096 and v <= 2040 supported");
  MPQC_T_CHECK(p->eps_occ && p->eps_vir && p->t1 && p->t2 && p->g_abij && p->g_aijk && p->g_abci,
               MPQC_T_ERR_BAD_ARG, "a tensor pointer is NULL");
  return MPQC_T_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------
extern "C" {

const char* mpqc_t_version(void) { return "mpqc_t_cuda 0.2 (sm_100a, abi 2)"; }

const char* mpqc_t_strerror(int status) {
  switch (status) {
    case MPQC_T_OK: return "ok";
    case MPQC_T_ERR_BAD_ARG: return "bad argument";
    case MPQC_T_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case MPQC_T_ERR_OOM: return "out of device memory";
    case MPQC_T_ERR_CUDA: return "CUDA runtime/driver error";
    case MPQC_T_ERR_NCCL: return "NCCL error";
    case MPQC_T_ERR_INTERNAL: return "internal error";
    default: return "unknown status";
  }
}

const char* mpqc_t_last_error(void) { return last_error_string().c_str(); }

int64_t mpqc_t_triple_count(int64_t o) {
  if (o < 1) return 0;
  return o * (o + 1) * (o + 2) / 6 - o;
}

int mpqc_t_triple_of_unit(int64_t o, int64_t unit, int32_t* i, int32_t* j, int32_t* k) {
  MPQC_T_CHECK(i && j && k, MPQC_T_ERR_BAD_ARG, "NULL output");
  MPQC_T_CHECK(o >= 1 && o <= 4096, MPQC_T_ERR_BAD_ARG, "need 1 <= o <= 4096");
  MPQC_T_CHECK(unit >= 0 && unit < mpqc_t_triple_count(o), MPQC_T_ERR_BAD_ARG, "unit out of range");
  const UnitIndex ux(o);
  int a, b, c;
  ux.triple(unit, a, b, c);
  *i = a;
  *j = b;
  *k = c;
  return MPQC_T_OK;
}

double mpqc_t_flops(int64_t o, int64_t v) { return 2.0 * (double)o * o * o * (double)v * v * v * (double)(v + o); }
double mpqc_t_unit_flops(int64_t o, int64_t v) { return 12.0 * (double)v * v * v * (double)(v + o); }

int mpqc_t_plan(int64_t o, int64_t v, int32_t flat, mpqc_t_plan_info* out) {
  MPQC_T_CHECK(out != nullptr, MPQC_T_ERR_BAD_ARG, "plan output is NULL");
  MPQC_T_CHECK(o >= 1 && v >= 1 && o <= 4096 && v <= 2040, MPQC_T_ERR_BAD_ARG, "need 1 <= o <= 4096, 1 <= v <= 2040");
  mpqc_t_handle h;                       // host-side fields only; no device call is made
  h.o = o;
  h.v = v;
  h.Kp = std::max<int64_t>(16, roundup(v + o, 8));
  h.flat = flat != 0;
  MPQC_T_TRY(plan(&h));
  memset(out, 0, sizeof(*out));
  out->kp = h.Kp;
  out->flat = h.flat;
  out->tp = h.tp;
  out->tq = h.tq;
  out->row_tiles = h.nmt;
  out->col_tiles = h.nnt;
  out->nfrag = h.nfrag;
  out->skip_last = h.skip_last;
  out->energy_tile_sets = h.ntt;
  const double mpad = (double)h.nmt * kBM, npad = (double)h.nnt * h.tn - 8.0 * h.skip_last;
  out->flop_efficiency = mpqc_t_unit_flops(o, v) / (3.0 * 2.0 * 2.0 * mpad * npad * (double)h.Kp);
  const double a = (double)o * v * v * h.Kp * 8.0;
  out->bytes_operands = a * (h.flat ? 2.0 : 1.0) + (double)o * o * v * h.Kp * 8.0 + (double)o * o * v * v * 8.0;
  return MPQC_T_OK;
}

int mpqc_t_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int mpqc_t_create(mpqc_t_handle** out, int64_t o, int64_t v, int32_t device) {
  MPQC_T_CHECK(out != nullptr, MPQC_T_ERR_BAD_ARG, "handle pointer is NULL");
  *out = nullptr;
  MPQC_T_CHECK(o >= 1 && v >= 1 && o <= 4096 && v <= 2040, MPQC_T_ERR_BAD_ARG, "need 1 <= o <= 4096, 1 <= v <= 2040");
  int ndev = mpqc_t_device_count();
  MPQC_T_CHECK(ndev > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible; the (T) path has no CPU fallback");
  MPQC_T_CHECK(device >= 0 && device < ndev, MPQC_T_ERR_BAD_ARG, "device ordinal out of range");
  MPQC_T_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MPQC_T_CUDA(cudaGetDeviceProperties(&prop, device));
  MPQC_T_CHECK(prop.major >= 10, MPQC_T_ERR_NO_DEVICE, "device is not sm_100-class (kernels are sm_100a only)");
  mpqc_t_handle* h = new (std::nothrow) mpqc_t_handle();
  MPQC_T_CHECK(h != nullptr, MPQC_T_ERR_OOM, "host allocation failed");
  h->device = device;
  h->o = o;
  h->v = v;
  h->Kp = std::max<int64_t>(16, roundup(v + o, 8));
  h->num_sms = prop.multiProcessorCount;
  int rc = [&]() -> int {
    MPQC_T_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // the big operand A (and its transposed copy) is allocated by the first upload: only then is it known whether
    // all o panels are resident (dense inputs; density-fitted inputs that fit) or a panel cache is used
    MPQC_T_CUDA(cudaMalloc(&h->B, (size_t)o * o * v * h->Kp * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->GV, (size_t)o * o * v * v * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->T1T, (size_t)o * v * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->eps_occ, (size_t)o * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->eps_vir, (size_t)v * sizeof(double)));
    MPQC_T_TRY(plan(h));
    std::vector<uint8_t> sets((size_t)h->ntt * 4);
    size_t n = 0;
    for (int a = 0; a < h->ntile; ++a)
      for (int b = 0; b <= a; ++b)
        for (int c = 0; c <= b; ++c) {
          sets[n++] = (uint8_t)a;
          sets[n++] = (uint8_t)b;
          sets[n++] = (uint8_t)c;
          sets[n++] = 0;
        }
    MPQC_T_CUDA(cudaMalloc(&h->tile_sets, sets.size()));
    MPQC_T_CUDA(cudaMemcpy(h->tile_sets, sets.data(), sets.size(), cudaMemcpyHostToDevice));
    GemmKernelFn fn = gemm_kernel_for(h->nfrag);
    MPQC_T_CHECK(fn != nullptr, MPQC_T_ERR_INTERNAL, "no W-contraction kernel for this column-fragment count");
    MPQC_T_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    MPQC_T_CUDA(cudaFuncSetAttribute(t_energy_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kEnergySmemBytes));
    return MPQC_T_OK;
  }();
  if (rc != MPQC_T_OK) {
    std::string keep = last_error_string();
    mpqc_t_destroy(h);
    last_error_string() = keep;
    return rc;
  }
  *out = h;
  return MPQC_T_OK;
}

int mpqc_t_destroy(mpqc_t_handle* h) {
  if (!h) return MPQC_T_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  free_work(h);
  cudaFree(h->A);
  cudaFree(h->AT);
  cudaFree(h->B);
  cudaFree(h->GV);
  cudaFree(h->T1T);
  cudaFree(h->eps_occ);
  cudaFree(h->eps_vir);
  cudaFree(h->tile_sets);
  cudaFree(h->triples_dev);
  cudaFree(h->unit_e_dev);
  cudaFree(h->arena);
  cudaFree(h->slot_map_dev);
  cudaFree(h->XaiT);
  cudaFree(h->XabT);
  cudaFree(h->T2raw);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
  return MPQC_T_OK;
}

void* mpqc_t_stream(mpqc_t_handle* h) { return h ? (void*)h->stream : nullptr; }

int mpqc_t_query(mpqc_t_handle* h, int32_t what, int64_t* value) {
  MPQC_T_CHECK(h != nullptr && value != nullptr, MPQC_T_ERR_BAD_ARG, "handle or output is NULL");
  switch (what) {
    case MPQC_T_QUERY_PANEL_SLOTS: *value = h->npanel; break;
    case MPQC_T_QUERY_PANEL_MODE: *value = h->panel_mode ? 1 : 0; break;
    case MPQC_T_QUERY_FLAT: *value = h->flat; break;
    case MPQC_T_QUERY_PANELS_BUILT: *value = h->panels_built; break;
    case MPQC_T_QUERY_PANEL_BLOCK: *value = h->panel_mode ? panel_block_edge(h) : h->o; break;
    default: return fail(MPQC_T_ERR_BAD_ARG, "unknown query", __FILE__, __LINE__);
  }
  return MPQC_T_OK;
}

int mpqc_t_set_df_block(mpqc_t_handle* h, int32_t block) {
  MPQC_T_CHECK(h != nullptr, MPQC_T_ERR_BAD_ARG, "handle is NULL");
  MPQC_T_CHECK(block >= -1 && block <= 4096, MPQC_T_ERR_BAD_ARG, "df block must be -1 (resident), 0 (automatic) or a block edge");
  h->df_block = block;
  return MPQC_T_OK;
}

int mpqc_t_plan_df(int64_t o, int64_t v, int64_t naux, int32_t block, int32_t flat, mpqc_t_df_plan_info* out) {
  MPQC_T_CHECK(out != nullptr, MPQC_T_ERR_BAD_ARG, "plan output is NULL");
  MPQC_T_CHECK(o >= 1 && v >= 1 && naux >= 1 && o <= 4096 && v <= 2040 && block >= 0, MPQC_T_ERR_BAD_ARG,
               "need 1 <= o <= 4096, 1 <= v <= 2040, naux >= 1, block >= 0");
  memset(out, 0, sizeof(*out));
  const double Kp = (double)std::max<int64_t>(16, roundup(v + o, 8)), Kx = (double)std::max<int64_t>(16, roundup(naux, 8));
  const int npanel = block > 0 ? (int)std::min<int64_t>(o, 3LL * block) : (int)o;
  out->npanel = npanel;
  out->panel_mode = npanel < o;
  out->block = out->panel_mode ? std::max(1, npanel / 3) : (int32_t)o;
  out->flat = flat != 0;
  out->bytes_panels = (double)npanel * v * v * Kp * 8.0 * (flat ? 2.0 : 1.0);
  out->bytes_b = (double)o * o * v * Kp * 8.0;
  out->bytes_gv = (double)o * o * v * v * 8.0;
  out->bytes_t2 = out->panel_mode ? (double)v * v * o * o * 8.0 : 0.0;
  out->bytes_factors = out->panel_mode ? ((double)o * v + (double)v * v) * Kx * 8.0 : 0.0;
  out->bytes_w_workspace = 3.0 * (double)v * v * (double)roundup(v, 16) * 8.0;
  out->bytes_total = out->bytes_panels + out->bytes_b + out->bytes_gv + out->bytes_t2 + out->bytes_factors + out->bytes_w_workspace;
  // panels built over the whole job ~ o^3 / (6 block^2) (one block of panels per occupied-block triple), 2 Kx v^3 FLOPs
  // each, against 2 o^3 v^3 (v+o) for the triples
  const double bo = (double)out->block;
  out->build_flop_fraction = out->panel_mode ? Kx / (6.0 * bo * bo * (double)(v + o))
                                             : (double)o * 2.0 * Kx * v * v * v / mpqc_t_flops(o, v);
  return MPQC_T_OK;
}

int mpqc_t_upload(mpqc_t_handle* h, const mpqc_t_problem* p, int32_t on_device, mpqc_t_stats* stats) {
  MPQC_T_CHECK(h != nullptr, MPQC_T_ERR_BAD_ARG, "handle is NULL");
  MPQC_T_TRY(validate_problem(p));
  MPQC_T_CHECK(p->o == h->o && p->v == h->v, MPQC_T_ERR_BAD_ARG, "problem dimensions differ from the handle's");
  MPQC_T_CUDA(cudaSetDevice(h->device));
  return upload_impl(h, p, on_device != 0, CommView(), stats);
}

static int run_range(mpqc_t_handle* h, int64_t first, int64_t stride, int64_t count, int32_t batch, double* partial_e,
                     double* unit_e, double* vblock_e, mpqc_t_stats* stats) {
  MPQC_T_CHECK(h != nullptr && partial_e != nullptr, MPQC_T_ERR_BAD_ARG, "handle or output is NULL");
  MPQC_T_CHECK(h->uploaded, MPQC_T_ERR_BAD_ARG, "mpqc_t_upload has not been called on this handle");
  if (stride <= 0) stride = 1;
  const UnitIndex ux(h->o);
  const int64_t nt = ux.count();
  MPQC_T_CHECK(first >= 0, MPQC_T_ERR_BAD_ARG, "unit_first < 0");
  int64_t avail = first < nt ? (nt - first + stride - 1) / stride : 0;
  if (count < 0 || count > avail) count = avail;
  *partial_e = 0.0;
  const double t0 = now_s();
  MPQC_T_CUDA(cudaSetDevice(h->device));
  DevBuf vb;
  if (vblock_e) {
    MPQC_T_TRY(vb.alloc((size_t)h->ntt));
    MPQC_T_CUDA(cudaMemsetAsync(vb.p, 0, (size_t)h->ntt * sizeof(double), h->stream));
  }
  if (count > 0) {
    std::vector<int64_t> units((size_t)count);
    for (int64_t u = 0; u < count; ++u) units[u] = first + u * stride;
    std::vector<double> ue((size_t)count);
    MPQC_T_TRY(run_units(h, ux, units.data(), count, batch, ue.data(), stats, getenv("MPQC_T_PROFILE") != nullptr, vb.p));
    double s = 0.0;
    for (int64_t u = 0; u < count; ++u) s += ue[u];   // fixed unit order -> deterministic
    *partial_e = s;
    if (unit_e) memcpy(unit_e, ue.data(), (size_t)count * sizeof(double));
  }
  if (vblock_e) {
    MPQC_T_CUDA(cudaMemcpyAsync(vblock_e, vb.p, (size_t)h->ntt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    MPQC_T_CUDA(cudaStreamSynchronize(h->stream));
  }
  if (stats) {
    stats->seconds_total += now_s() - t0;
    stats->ngpu = 1;
  }
  return MPQC_T_OK;
}

int mpqc_t_run(mpqc_t_handle* h, int64_t first, int64_t stride, int64_t count, int32_t batch, double* partial_e,
               double* unit_e, mpqc_t_stats* stats) {
  return run_range(h, first, stride, count, batch, partial_e, unit_e, nullptr, stats);
}

int mpqc_t_run_comm(mpqc_t_handle* h, mpqc_t_comm* c, int64_t first, int64_t stride, int64_t count, int32_t batch,
                    double* total_e, double* unit_e, mpqc_t_stats* stats) {
  MPQC_T_CHECK(h != nullptr && c != nullptr && total_e != nullptr, MPQC_T_ERR_BAD_ARG, "handle, communicator or output is NULL");
  MPQC_T_CHECK(h->uploaded, MPQC_T_ERR_BAD_ARG, "mpqc_t_upload has not been called on this handle");
  MPQC_T_CHECK(c->members.size() == 1 && c->members[0].device == h->device, MPQC_T_ERR_BAD_ARG,
               "mpqc_t_run_comm needs a rank-mode communicator whose device is the handle's");
  if (stride <= 0) stride = 1;
  const UnitIndex ux(h->o);
  const int64_t nt = ux.count();
  MPQC_T_CHECK(first >= 0, MPQC_T_ERR_BAD_ARG, "unit_first < 0");
  const int64_t avail = first < nt ? (nt - first + stride - 1) / stride : 0;
  if (count < 0 || count > avail) count = avail;
  const CommMember& m = c->members[0];
  const int R = c->nranks, r = m.rank;
  *total_e = 0.0;
  const double t0 = now_s();
  MPQC_T_CUDA(cudaSetDevice(h->device));
  // this rank's share: job positions r, r+R, ...
  const int64_t mine = count > r ? (count - r + R - 1) / R : 0;
  std::vector<int64_t> units((size_t)mine);
  for (int64_t q = 0; q < mine; ++q) units[(size_t)q] = first + (r + q * R) * stride;
  std::vector<double> ue_mine((size_t)mine);
  int rc = run_units(h, ux, units.data(), mine, batch, ue_mine.data(), stats, getenv("MPQC_T_PROFILE") != nullptr);
  // Job-length vector on the device: own units at their job positions (strided device copy from the per-unit results
  // run_units left in h->unit_e_dev), zeros elsewhere; one ncclAllReduce on the handle's stream completes it on every
  // rank (x + 0 + ... + 0 is exact).  Reached also after a local failure, with a raised status word, so that no peer
  // blocks (the vector lives in the communicator's pre-allocated scratch when it fits).
  std::vector<double> all((size_t)count + 1, 0.0);
  if (R > 1) {
    const NcclApi& nc = nccl_api();
    const bool fits = (size_t)count + 1 <= kCommScratchDoubles;
    if (fits) {
      int r2 = cuda_status(cudaMemsetAsync(m.scratch, 0, ((size_t)count + 1) * sizeof(double), h->stream), "memset", __FILE__, __LINE__);
      if (rc == MPQC_T_OK && r2 == MPQC_T_OK && mine > 0)
        r2 = cuda_status(cudaMemcpy2DAsync(m.scratch + r, (size_t)R * sizeof(double), h->unit_e_dev, sizeof(double),
                                           sizeof(double), (size_t)mine, cudaMemcpyDeviceToDevice, h->stream),
                         "cudaMemcpy2DAsync(unit energies)", __FILE__, __LINE__);
      if (rc != MPQC_T_OK || r2 != MPQC_T_OK) {
        const double one = 1.0;
        cudaMemcpyAsync(m.scratch + count, &one, sizeof(double), cudaMemcpyHostToDevice, h->stream);
      }
      const std::string keep = last_error_string();
      int r3 = nccl_status(nc.AllReduce(m.scratch, m.scratch, (size_t)count + 1, kNcclFloat64, kNcclSum, m.comm, h->stream),
                           "ncclAllReduce(unit energies)", __FILE__, __LINE__);
      if (r3 == MPQC_T_OK)
        r3 = cuda_status(cudaMemcpyAsync(all.data(), m.scratch, ((size_t)count + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream),
                         "cudaMemcpyAsync(summed unit energies)", __FILE__, __LINE__);
      if (r3 == MPQC_T_OK) r3 = cuda_status(cudaStreamSynchronize(h->stream), "cudaStreamSynchronize", __FILE__, __LINE__);
      if (rc != MPQC_T_OK) last_error_string() = keep;
      if (rc == MPQC_T_OK) rc = r2 != MPQC_T_OK ? r2 : r3;
    } else {
      for (int64_t q = 0; q < mine && rc == MPQC_T_OK; ++q) all[(size_t)(r + q * R)] = ue_mine[(size_t)q];
      all[(size_t)count] = rc == MPQC_T_OK ? 0.0 : 1.0;
      CommView cv;
      cv.rank = r;
      cv.nranks = R;
      cv.comm = m.comm;
      cv.scratch = m.scratch;
      const std::string keep = last_error_string();
      const int r3 = allreduce_host_vector(cv, all.data(), all.size(), h->stream);
      if (rc != MPQC_T_OK) last_error_string() = keep;
      if (rc == MPQC_T_OK) rc = r3;
    }
    if (rc == MPQC_T_OK && all[(size_t)count] > 0.5)
      rc = fail(MPQC_T_ERR_INTERNAL, "another rank of the (T) communicator failed during the triples loop", __FILE__, __LINE__);
    if (stats) {
      stats->bytes_d2h += ((int64_t)count + 1) * 8;
      stats->kernel_launches += 1;   // the all-reduce
    }
  } else {
    for (int64_t q = 0; q < mine; ++q) all[(size_t)q] = ue_mine[(size_t)q];
  }
  MPQC_T_TRY(rc);
  double s = 0.0;
  for (int64_t u = 0; u < count; ++u) s += all[(size_t)u];   // job order: identical on every rank and for every R
  *total_e = s;
  if (unit_e) memcpy(unit_e, all.data(), (size_t)count * sizeof(double));
  if (stats) {
    stats->seconds_total += now_s() - t0;
    stats->ngpu = 1;
  }
  return MPQC_T_OK;
}

int mpqc_t_run_vblocks(mpqc_t_handle* h, int64_t first, int64_t stride, int64_t count, int32_t batch, double* partial_e,
                       double* unit_e, double* vblock_e, mpqc_t_stats* stats) {
  MPQC_T_CHECK(vblock_e != nullptr, MPQC_T_ERR_BAD_ARG, "vblock_e is NULL");
  return run_range(h, first, stride, count, batch, partial_e, unit_e, vblock_e, stats);
}

int mpqc_t_debug_w(mpqc_t_handle* h, int32_t i, int32_t j, int32_t k, double* w_host) {
  MPQC_T_CHECK(h && w_host, MPQC_T_ERR_BAD_ARG, "NULL argument");
  MPQC_T_CHECK(h->uploaded, MPQC_T_ERR_BAD_ARG, "mpqc_t_upload has not been called on this handle");
  MPQC_T_CHECK(i >= 0 && j >= 0 && k >= 0 && i < h->o && j < h->o && k < h->o, MPQC_T_ERR_BAD_ARG, "triple out of range");
  MPQC_T_CUDA(cudaSetDevice(h->device));
  MPQC_T_TRY(ensure_work(h, 1));
  MPQC_T_TRY(ensure_units(h, 1));
  int tri[3] = {i, j, k};
  MPQC_T_CUDA(cudaMemcpyAsync(h->triples_dev, tri, sizeof(tri), cudaMemcpyHostToDevice, h->stream));
  if (h->panel_mode) {
    std::vector<int> need = {i, j, k};
    std::sort(need.begin(), need.end());
    need.erase(std::unique(need.begin(), need.end()), need.end());
    MPQC_T_TRY(ensure_panels(h, need, nullptr));
  }
  MPQC_T_TRY(launch_gemm(h, 1, h->triples_dev));
  const int64_t v = h->v, ldw = h->ldw;
  std::vector<double> n((size_t)3 * v * v * ldw);
  MPQC_T_CUDA(cudaMemcpyAsync(n.data(), h->W, n.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  MPQC_T_CUDA(cudaStreamSynchronize(h->stream));
  const double* n0 = n.data();
  const double* n1 = n0 + v * v * ldw;
  const double* n2 = n1 + v * v * ldw;
  for (int64_t a = 0; a < v; ++a)
    for (int64_t b = 0; b < v; ++b)
      for (int64_t c = 0; c < v; ++c)
        w_host[(a * v + b) * v + c] = n0[(a * v + b) * ldw + c] + n1[(a * v + c) * ldw + b] + n2[(c * v + b) * ldw + a];
  return MPQC_T_OK;
}

int mpqc_t_w_batch(mpqc_t_handle* h, const int32_t* triples, int64_t n, double* w_out, int32_t out_on_device) {
  MPQC_T_CHECK(h && triples && w_out, MPQC_T_ERR_BAD_ARG, "NULL argument");
  MPQC_T_CHECK(h->uploaded, MPQC_T_ERR_BAD_ARG, "mpqc_t_upload has not been called on this handle");
  MPQC_T_CHECK(n >= 0, MPQC_T_ERR_BAD_ARG, "n < 0");
  for (int64_t q = 0; q < 3 * n; ++q)
    MPQC_T_CHECK(triples[q] >= 0 && triples[q] < h->o, MPQC_T_ERR_BAD_ARG, "occupied index out of range");
  if (n == 0) return MPQC_T_OK;
  MPQC_T_CUDA(cudaSetDevice(h->device));
  const int64_t v = h->v, v3 = v * v * v;
  int batch = (int)std::min<int64_t>(std::max(1, auto_batch(h)), n);
  if (h->panel_mode) batch = std::min(batch, std::max(1, h->npanel / 3));   // a batch never needs more panels than the pool holds
  MPQC_T_TRY(ensure_work(h, batch));
  MPQC_T_TRY(ensure_units(h, n));
  MPQC_T_CUDA(cudaMemcpyAsync(h->triples_dev, triples, (size_t)n * 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  DevBuf stage;                       // host output: assembled on the device batch by batch, then copied back
  if (!out_on_device) MPQC_T_TRY(stage.alloc((size_t)batch * v3));
  for (int64_t off = 0; off < n; off += batch) {
    const int nb = (int)std::min<int64_t>(batch, n - off);
    if (h->panel_mode) {
      std::vector<int> need(triples + 3 * off, triples + 3 * (off + nb));
      std::sort(need.begin(), need.end());
      need.erase(std::unique(need.begin(), need.end()), need.end());
      MPQC_T_TRY(ensure_panels(h, need, nullptr));
    }
    MPQC_T_TRY(launch_gemm(h, nb, h->triples_dev + 3 * off));
    double* dst = out_on_device ? w_out + off * v3 : stage.p;
    w_assemble_kernel<<<dim3((unsigned)(h->ntile * h->ntile * h->ntile), (unsigned)nb), 512, 0, h->stream>>>(h->W, dst, (int)v,
                                                                                                         h->ldw, h->ntile);
    MPQC_T_CUDA(cudaGetLastError());
    if (!out_on_device)
      MPQC_T_CUDA(cudaMemcpyAsync(w_out + off * v3, stage.p, (size_t)nb * v3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  MPQC_T_CUDA(cudaStreamSynchronize(h->stream));
  MPQC_T_CUDA(cudaGetLastError());
  return MPQC_T_OK;
}

}  // extern "C"

namespace {

typedef std::function<int(mpqc_t_handle*, const CommView&, mpqc_t_stats*)> UploadFn;

// Shared driver of the one-shot entry points.  The job is the unit list  first, first+stride, ... (count of them);
// it is sharded over the W workers of the communicator (worker w takes job positions w, w+W, ...; when every worker
// lives in this process the last 1/8 is handed out by an atomic counter instead -- work stealing), each worker
// uploads/replicates the inputs onto its GPU, runs its units, and the per-unit energies are summed: on the host
// when no communicator is involved, else by one ncclAllReduce over the unit-energy vector.
int energy_impl(const int64_t prob_o, const int64_t prob_v, const UploadFn& upload, const mpqc_t_options* opt_in,
                mpqc_t_comm* comm_in, double* e_t, mpqc_t_stats* stats_out) {
  mpqc_t_options opt;
  memset(&opt, 0, sizeof(opt));
  if (opt_in) opt = *opt_in;
  const int ndev = mpqc_t_device_count();
  MPQC_T_CHECK(ndev > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible; the (T) path has no CPU fallback");

  // ---- who works: the communicator's members that live here, or opt.ngpu devices without a communicator ----
  struct TempComm {
    mpqc_t_comm* c = nullptr;
    ~TempComm() { mpqc_t_comm_destroy(c); }
  } temp;
  mpqc_t_comm* comm = comm_in;
  if (!comm && opt.use_nccl && opt.ngpu > 1) {
    // one-shot convenience: a communicator that lives for this call only (its set-up costs seconds; hosts that call
    // more than once, or that can prepare ahead of time, should hold a persistent mpqc_t_comm)
    MPQC_T_TRY(mpqc_t_comm_create_local(&temp.c, opt.ngpu, opt.device_ids));
    comm = temp.c;
  }
  std::vector<CommView> views;
  std::vector<int> devs;
  int nranks = 1;
  if (comm) {
    nranks = comm->nranks;
    for (const CommMember& m : comm->members) {
      CommView cv;
      cv.rank = m.rank;
      cv.nranks = comm->nranks;
      cv.comm = m.comm;
      cv.scratch = m.scratch;
      views.push_back(cv);
      devs.push_back(m.device);
    }
  } else {
    const int ngpu = opt.ngpu > 0 ? opt.ngpu : 1;
    for (int g = 0; g < ngpu; ++g) {
      const int d = opt.device_ids ? opt.device_ids[g] : g;
      MPQC_T_CHECK(d >= 0 && d < ndev, MPQC_T_ERR_BAD_ARG, "device ordinal out of range");
      views.push_back(CommView());
      devs.push_back(d);
    }
  }
  const int nlocal = (int)views.size();
  const bool exchange = comm != nullptr && nranks > 1;
  MPQC_T_CHECK(!(opt.inputs_on_device && nlocal != 1), MPQC_T_ERR_BAD_ARG,
               "inputs_on_device requires one device per process");
  const double t0 = now_s();
  mpqc_t_stats stats;
  memset(&stats, 0, sizeof(stats));

  // ---- the job and its split ----
  const UnitIndex ux(prob_o);
  const int64_t nt = ux.count();
  const int64_t stride = opt.unit_stride > 0 ? opt.unit_stride : 1;
  MPQC_T_CHECK(opt.unit_first >= 0, MPQC_T_ERR_BAD_ARG, "unit_first < 0");
  const int64_t avail = opt.unit_first < nt ? (nt - opt.unit_first + stride - 1) / stride : 0;
  const int64_t count = (opt.unit_count < 0 || opt.unit_count > avail) ? avail : opt.unit_count;
  std::vector<double> unit_e((size_t)count, 0.0);   // each slot is written by exactly one worker thread
  const int W = comm ? nranks : nlocal;              // workers over which the job is split
  const bool all_local = !comm || comm->local;       // work stealing needs shared memory
  const int64_t static_n = (W > 1 && all_local) ? (count / 8) * 7 / W * W : count;
  std::atomic<int64_t> tail_next(static_n);
  const bool profile = getenv("MPQC_T_PROFILE") != nullptr;

  std::vector<mpqc_t_stats> gstats(nlocal);
  std::vector<int> rcs(nlocal, MPQC_T_OK);
  std::vector<std::string> msgs(nlocal);
  std::vector<double> reduced;                       // unit energies after the all-reduce (written by local worker 0)
  if (exchange) reduced.assign((size_t)count, 0.0);

  auto worker = [&](int g) {
    mpqc_t_stats& gs = gstats[g];
    memset(&gs, 0, sizeof(gs));
    const CommView& cv = views[g];
    const int wrank = comm ? cv.rank : g;
    mpqc_t_handle* h = nullptr;
    const double tw0 = now_s();
    if (comm && (int)comm->cached.size() > g && comm->cached[(size_t)g]) {
      // device memory of the previous call on this member: re-used when the problem has the same shape
      mpqc_t_handle* c = comm->cached[(size_t)g];
      comm->cached[(size_t)g] = nullptr;
      if (c->o == prob_o && c->v == prob_v && c->device == devs[g]) h = c;
      else mpqc_t_destroy(c);
    }
    int rc = h ? MPQC_T_OK : mpqc_t_create(&h, prob_o, prob_v, devs[g]);
    if (rc == MPQC_T_OK) h->df_block = opt.df_block;
    cudaStream_t cst = nullptr;                       // stream of this worker's collectives
    if (exchange) {
      cudaSetDevice(devs[g]);
      if (h) cst = h->stream;
      else if (cudaStreamCreateWithFlags(&cst, cudaStreamNonBlocking) != cudaSuccess) cst = nullptr;
      // agreement #1: every rank holds its operand memory, or nobody starts the replicated upload
      const int nfail = cst ? count_failed_ranks(cv, rc, cst) : -1;
      if (rc == MPQC_T_OK && nfail != 0)
        rc = fail(nfail < 0 ? MPQC_T_ERR_NCCL : MPQC_T_ERR_INTERNAL,
                  "another rank of the (T) communicator failed to set up its device", __FILE__, __LINE__);
    }
    const double tw1 = now_s();
    if (rc == MPQC_T_OK) rc = upload(h, cv, &gs);
    const double tw2 = now_s();
    std::vector<int64_t> done_idx;     // job positions this worker produced
    if (rc == MPQC_T_OK && h->panel_mode) {
      // Panel-cache mode: shard by occupied-block triple, not by unit -- a worker that holds a group's panels runs
      // the whole group.  Groups are dealt largest-first to the least loaded worker (same answer on every rank).
      const int bo = panel_block_edge(h);
      std::vector<std::pair<int64_t, int64_t>> keyed((size_t)count);   // (group key, job position)
      for (int64_t q = 0; q < count; ++q) {
        int i, j, k;
        ux.triple(opt.unit_first + q * stride, i, j, k);
        keyed[(size_t)q] = std::make_pair(block_key(i, j, k, bo), q);
      }
      std::sort(keyed.begin(), keyed.end());
      std::vector<std::pair<int64_t, int64_t>> groups;                  // (size, first index into keyed)
      for (int64_t a = 0; a < count;) {
        int64_t b = a;
        while (b < count && keyed[(size_t)b].first == keyed[(size_t)a].first) ++b;
        groups.push_back(std::make_pair(b - a, a));
        a = b;
      }
      std::stable_sort(groups.begin(), groups.end(),
                       [](const std::pair<int64_t, int64_t>& x, const std::pair<int64_t, int64_t>& y) { return x.first > y.first; });
      std::vector<int64_t> load((size_t)W, 0), mine, idx;
      for (const auto& gsz : groups) {
        const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        load[(size_t)w] += gsz.first;
        if (w == wrank)
          for (int64_t a = gsz.second; a < gsz.second + gsz.first; ++a) mine.push_back(keyed[(size_t)a].second);
      }
      idx.resize(mine.size());
      std::vector<double> e(mine.size());
      for (size_t q = 0; q < mine.size(); ++q) idx[q] = opt.unit_first + mine[q] * stride;
      rc = run_units(h, ux, idx.data(), (int64_t)idx.size(), opt.batch, e.data(), &gs, profile);
      if (rc == MPQC_T_OK)
        for (size_t q = 0; q < mine.size(); ++q) {
          unit_e[(size_t)mine[q]] = e[q];
          done_idx.push_back(mine[q]);
        }
    } else if (rc == MPQC_T_OK) {
      // static share
      std::vector<int64_t> mine, idx;
      for (int64_t q = wrank; q < static_n; q += W) mine.push_back(q);
      idx.resize(mine.size());
      std::vector<double> e(mine.size());
      for (size_t q = 0; q < mine.size(); ++q) idx[q] = opt.unit_first + mine[q] * stride;
      rc = run_units(h, ux, idx.data(), (int64_t)idx.size(), opt.batch, e.data(), &gs, profile);
      if (rc == MPQC_T_OK)
        for (size_t q = 0; q < mine.size(); ++q) {
          unit_e[(size_t)mine[q]] = e[q];
          done_idx.push_back(mine[q]);
        }
      // work-stealing tail (only when all workers share this process)
      const int64_t chunk = opt.steal_chunk > 0 ? opt.steal_chunk : std::max<int64_t>(1, auto_batch(h)) * 4;
      while (rc == MPQC_T_OK && static_n < count) {
        const int64_t s0 = tail_next.fetch_add(chunk);
        if (s0 >= count) break;
        const int64_t n = std::min(chunk, count - s0);
        std::vector<int64_t> ids((size_t)n);
        for (int64_t q = 0; q < n; ++q) ids[(size_t)q] = opt.unit_first + (s0 + q) * stride;
        std::vector<double> e2((size_t)n);
        rc = run_units(h, ux, ids.data(), n, opt.batch, e2.data(), &gs, profile);
        if (rc == MPQC_T_OK)
          for (int64_t q = 0; q < n; ++q) {
            unit_e[(size_t)(s0 + q)] = e2[(size_t)q];
            done_idx.push_back(s0 + q);
          }
      }
    }
    if (exchange && cst) {
      // the path's one arithmetic collective (replaces gop.sum, ccsd_t.h:692).  Every rank reaches it, also after
      // a local failure (it then contributes zeros and a raised status word).
      cudaSetDevice(devs[g]);
      std::vector<double> mine((size_t)count + 1, 0.0);
      if (rc == MPQC_T_OK)
        for (int64_t q : done_idx) mine[(size_t)q] = unit_e[(size_t)q];   // only the slots this worker wrote itself
      mine[(size_t)count] = rc == MPQC_T_OK ? 0.0 : 1.0;
      const std::string keep = last_error_string();
      const int r2 = allreduce_host_vector(cv, mine.data(), mine.size(), cst);
      if (rc != MPQC_T_OK) last_error_string() = keep;
      if (rc == MPQC_T_OK) {
        if (r2 != MPQC_T_OK) rc = r2;
        else if (mine[(size_t)count] > 0.5)
          rc = fail(MPQC_T_ERR_INTERNAL, "another rank of the (T) communicator failed during the triples loop", __FILE__, __LINE__);
        else if (g == 0) std::copy(mine.begin(), mine.begin() + count, reduced.begin());
      }
      gs.bytes_h2d += (int64_t)mine.size() * 8;
      gs.bytes_d2h += (int64_t)mine.size() * 8;
    }
    if (rc != MPQC_T_OK) msgs[g] = last_error_string();
    const double tw3 = now_s();
    if (!h && cst) cudaStreamDestroy(cst);
    if (comm && rc == MPQC_T_OK && (int)comm->cached.size() > g) comm->cached[(size_t)g] = h;   // keep the memory for the next call
    else mpqc_t_destroy(h);
    if (opt.verbose >= 2)
      printf("  [mpqc_t] rank %d gpu %d: create %.3f s, upload+relayout %.3f s, triples+sum %.3f s, destroy %.3f s\n", wrank,
             devs[g], tw1 - tw0, tw2 - tw1, tw3 - tw2, now_s() - tw3);
    rcs[g] = rc;
  };

  if (nlocal == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int g = 0; g < nlocal; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();   // joined before returning (SURVEY 8b threading contract)
  }
  for (int g = 0; g < nlocal; ++g)
    if (rcs[g] != MPQC_T_OK) {
      last_error_string() = msgs[g];
      return rcs[g];
    }
  double e = 0.0;
  const std::vector<double>& final_e = exchange ? reduced : unit_e;
  for (int64_t u = 0; u < count; ++u) e += final_e[(size_t)u];   // unit order: bit-identical for any number of GPUs
  *e_t = e;

  for (int g = 0; g < nlocal; ++g) {
    stats.seconds_upload = std::max(stats.seconds_upload, gstats[g].seconds_upload);
    stats.seconds_relayout = std::max(stats.seconds_relayout, gstats[g].seconds_relayout);
    stats.seconds_compute = std::max(stats.seconds_compute, gstats[g].seconds_compute);
    stats.seconds_contract = std::max(stats.seconds_contract, gstats[g].seconds_contract);
    stats.seconds_energy = std::max(stats.seconds_energy, gstats[g].seconds_energy);
    stats.flops += gstats[g].flops;
    stats.flops_executed += gstats[g].flops_executed;
    stats.units += gstats[g].units;
    stats.kernel_launches += gstats[g].kernel_launches;
    stats.bytes_h2d += gstats[g].bytes_h2d;
    stats.bytes_d2h += gstats[g].bytes_d2h;
  }
  stats.ngpu = nlocal;
  stats.seconds_total = now_s() - t0;
  if (opt.verbose) {
    // same line the reference prints, ccsd_t.h:175
    printf("(T) Energy: %.15g Time: %g S\n", e, stats.seconds_total);
    fflush(stdout);
  }
  if (stats_out) *stats_out = stats;
  return MPQC_T_OK;
}

int validate_df_problem(const mpqc_t_df_problem* p) {
  MPQC_T_CHECK(p != nullptr, MPQC_T_ERR_BAD_ARG, "problem is NULL");
  MPQC_T_CHECK(p->o >= 1 && p->v >= 1 && p->naux >= 1, MPQC_T_ERR_BAD_ARG, "o, v and naux must be >= 1");
  MPQC_T_CHECK(p->o <= 4096 && p->v <= 2040 && p->naux <= (1 << 20), MPQC_T_ERR_BAD_ARG,
               "o <= 4096, v <= 2040, naux <= 2^20 supported");
  MPQC_T_CHECK(p->eps_occ && p->eps_vir && p->t1 && p->t2 && p->x_ab && p->x_ij && p->x_ai, MPQC_T_ERR_BAD_ARG,
               "a tensor pointer is NULL");
  return MPQC_T_OK;
}

}  // namespace

extern "C" {

static int energy_dense(mpqc_t_comm* c, const mpqc_t_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats) {
  MPQC_T_CHECK(e_t != nullptr, MPQC_T_ERR_BAD_ARG, "e_t is NULL");
  MPQC_T_TRY(validate_problem(p));
  const bool on_dev = opt && opt->inputs_on_device;
  return energy_impl(p->o, p->v,
                     [&](mpqc_t_handle* h, const CommView& cv, mpqc_t_stats* st) {
                       MPQC_T_CUDA(cudaSetDevice(h->device));
                       return upload_impl(h, p, on_dev, cv, st);
                     },
                     opt, c, e_t, stats);
}

static int energy_df(mpqc_t_comm* c, const mpqc_t_df_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats) {
  MPQC_T_CHECK(e_t != nullptr, MPQC_T_ERR_BAD_ARG, "e_t is NULL");
  MPQC_T_TRY(validate_df_problem(p));
  const bool on_dev = opt && opt->inputs_on_device;
  return energy_impl(p->o, p->v,
                     [&](mpqc_t_handle* h, const CommView& cv, mpqc_t_stats* st) {
                       MPQC_T_CUDA(cudaSetDevice(h->device));
                       return upload_df_impl(h, p, on_dev, cv, st);
                     },
                     opt, c, e_t, stats);
}

int mpqc_t_energy(const mpqc_t_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats) {
  return energy_dense(nullptr, p, opt, e_t, stats);
}

int mpqc_t_energy_df(const mpqc_t_df_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats) {
  return energy_df(nullptr, p, opt, e_t, stats);
}

int mpqc_t_energy_comm(mpqc_t_comm* c, const mpqc_t_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats) {
  MPQC_T_CHECK(c != nullptr, MPQC_T_ERR_BAD_ARG, "communicator is NULL");
  return energy_dense(c, p, opt, e_t, stats);
}

int mpqc_t_energy_df_comm(mpqc_t_comm* c, const mpqc_t_df_problem* p, const mpqc_t_options* opt, double* e_t,
                          mpqc_t_stats* stats) {
  MPQC_T_CHECK(c != nullptr, MPQC_T_ERR_BAD_ARG, "communicator is NULL");
  return energy_df(c, p, opt, e_t, stats);
}

// ---- communicator -------------------------------------------------------------------------------------------
int mpqc_t_comm_unique_id(mpqc_t_unique_id* id) {
  MPQC_T_CHECK(id != nullptr, MPQC_T_ERR_BAD_ARG, "id is NULL");
  const NcclApi& nc = nccl_api();
  MPQC_T_CHECK(nc.ok, MPQC_T_ERR_NCCL, "libnccl.so.2 could not be loaded");
  MPQC_T_NCCL(nc.GetUniqueId(reinterpret_cast<NcclUniqueId*>(id)));
  return MPQC_T_OK;
}

static int comm_member_init(CommMember& m) {
  MPQC_T_CUDA(cudaSetDevice(m.device));
  MPQC_T_CUDA(cudaMalloc(&m.scratch, kCommScratchDoubles * sizeof(double)));
  return MPQC_T_OK;
}

int mpqc_t_comm_create_rank(mpqc_t_comm** out, int32_t nranks, int32_t rank, const mpqc_t_unique_id* id, int32_t device) {
  MPQC_T_CHECK(out != nullptr, MPQC_T_ERR_BAD_ARG, "communicator pointer is NULL");
  *out = nullptr;
  MPQC_T_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, MPQC_T_ERR_BAD_ARG, "need 0 <= rank < nranks");
  MPQC_T_CHECK(id != nullptr || nranks == 1, MPQC_T_ERR_BAD_ARG, "unique id is NULL");
  const int ndev = mpqc_t_device_count();
  MPQC_T_CHECK(ndev > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible; the (T) path has no CPU fallback");
  MPQC_T_CHECK(device >= 0 && device < ndev, MPQC_T_ERR_BAD_ARG, "device ordinal out of range");
  mpqc_t_comm* c = new (std::nothrow) mpqc_t_comm();
  MPQC_T_CHECK(c != nullptr, MPQC_T_ERR_OOM, "host allocation failed");
  c->nranks = nranks;
  c->local = nranks == 1;
  c->members.resize(1);
  c->cached.assign(1, nullptr);
  c->members[0].rank = rank;
  c->members[0].device = device;
  int rc = comm_member_init(c->members[0]);
  if (rc == MPQC_T_OK && nranks > 1) {
    const NcclApi& nc = nccl_api();
    if (!nc.ok) rc = fail(MPQC_T_ERR_NCCL, "libnccl.so.2 could not be loaded", __FILE__, __LINE__);
    else {
      NcclUniqueId uid;
      memcpy(&uid, id, sizeof(uid));
      rc = nccl_status(nc.CommInitRank(&c->members[0].comm, nranks, uid, rank), "ncclCommInitRank", __FILE__, __LINE__);
    }
  }
  if (rc != MPQC_T_OK) {
    const std::string keep = last_error_string();
    mpqc_t_comm_destroy(c);
    last_error_string() = keep;
    return rc;
  }
  *out = c;
  return MPQC_T_OK;
}

int mpqc_t_comm_create_local(mpqc_t_comm** out, int32_t ngpu, const int32_t* device_ids) {
  MPQC_T_CHECK(out != nullptr, MPQC_T_ERR_BAD_ARG, "communicator pointer is NULL");
  *out = nullptr;
  MPQC_T_CHECK(ngpu >= 1, MPQC_T_ERR_BAD_ARG, "ngpu must be >= 1");
  const int ndev = mpqc_t_device_count();
  MPQC_T_CHECK(ndev > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible; the (T) path has no CPU fallback");
  std::vector<int> devs(ngpu);
  for (int g = 0; g < ngpu; ++g) {
    devs[g] = device_ids ? device_ids[g] : g;
    MPQC_T_CHECK(devs[g] >= 0 && devs[g] < ndev, MPQC_T_ERR_BAD_ARG, "device ordinal out of range");
  }
  mpqc_t_comm* c = new (std::nothrow) mpqc_t_comm();
  MPQC_T_CHECK(c != nullptr, MPQC_T_ERR_OOM, "host allocation failed");
  c->nranks = ngpu;
  c->local = true;
  c->members.resize(ngpu);
  c->cached.assign((size_t)ngpu, nullptr);
  // CUDA contexts are created here, one thread per device, so that the serial seconds of context creation in a
  // process that drives 8 GPUs are paid once and in parallel, outside the (T) call
  std::vector<int> rcs(ngpu, MPQC_T_OK);
  std::vector<std::string> msgs(ngpu);
  {
    std::vector<std::thread> th;
    for (int g = 0; g < ngpu; ++g)
      th.emplace_back([&, g] {
        c->members[g].rank = g;
        c->members[g].device = devs[g];
        rcs[g] = comm_member_init(c->members[g]);
        if (rcs[g] != MPQC_T_OK) msgs[g] = last_error_string();
      });
    for (auto& t : th) t.join();
  }
  int rc = MPQC_T_OK;
  for (int g = 0; g < ngpu && rc == MPQC_T_OK; ++g)
    if (rcs[g] != MPQC_T_OK) {
      last_error_string() = msgs[g];
      rc = rcs[g];
    }
  if (rc == MPQC_T_OK && ngpu > 1) {
    const NcclApi& nc = nccl_api();
    if (!nc.ok) rc = fail(MPQC_T_ERR_NCCL, "libnccl.so.2 could not be loaded", __FILE__, __LINE__);
    else {
      std::vector<NcclApi::comm_t> comms(ngpu, nullptr);
      rc = nccl_status(nc.CommInitAll(comms.data(), ngpu, devs.data()), "ncclCommInitAll", __FILE__, __LINE__);
      for (int g = 0; g < ngpu; ++g) c->members[g].comm = comms[g];
    }
  }
  if (rc != MPQC_T_OK) {
    const std::string keep = last_error_string();
    mpqc_t_comm_destroy(c);
    last_error_string() = keep;
    return rc;
  }
  *out = c;
  return MPQC_T_OK;
}

int mpqc_t_comm_release_cache(mpqc_t_comm* c) {
  if (!c) return MPQC_T_OK;
  for (mpqc_t_handle*& h : c->cached) {
    mpqc_t_destroy(h);
    h = nullptr;
  }
  return MPQC_T_OK;
}

int mpqc_t_comm_destroy(mpqc_t_comm* c) {
  if (!c) return MPQC_T_OK;
  mpqc_t_comm_release_cache(c);
  const NcclApi& nc = nccl_api();
  for (CommMember& m : c->members) {
    cudaSetDevice(m.device);
    if (m.comm && nc.ok) nc.CommDestroy(m.comm);
    cudaFree(m.scratch);
  }
  cudaGetLastError();
  delete c;
  return MPQC_T_OK;
}

int mpqc_t_comm_size(const mpqc_t_comm* c) { return c ? c->nranks : 0; }

int mpqc_t_host_alloc(void** ptr, size_t bytes) {
  MPQC_T_CHECK(ptr != nullptr, MPQC_T_ERR_BAD_ARG, "pointer is NULL");
  *ptr = nullptr;
  MPQC_T_CHECK(mpqc_t_device_count() > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible; the (T) path has no CPU fallback");
  MPQC_T_CUDA(cudaHostAlloc(ptr, std::max<size_t>(bytes, 1), cudaHostAllocPortable));
  return MPQC_T_OK;
}

int mpqc_t_host_free(void* ptr) {
  if (ptr) MPQC_T_CUDA(cudaFreeHost(ptr));
  return MPQC_T_OK;
}

int mpqc_t_upload_df(mpqc_t_handle* h, const mpqc_t_df_problem* p, int32_t on_device, mpqc_t_stats* stats) {
  MPQC_T_CHECK(h != nullptr, MPQC_T_ERR_BAD_ARG, "handle is NULL");
  MPQC_T_TRY(validate_df_problem(p));
  MPQC_T_CHECK(p->o == h->o && p->v == h->v, MPQC_T_ERR_BAD_ARG, "problem dimensions differ from the handle's");
  MPQC_T_CUDA(cudaSetDevice(h->device));
  return upload_df_impl(h, p, on_device != 0, CommView(), stats);
}

int mpqc_t_microbench(int32_t device, int32_t which, double* tflops) {
  MPQC_T_CHECK(tflops != nullptr, MPQC_T_ERR_BAD_ARG, "tflops is NULL");
  int ndev = mpqc_t_device_count();
  MPQC_T_CHECK(ndev > 0, MPQC_T_ERR_NO_DEVICE, "no CUDA device visible");
  MPQC_T_CHECK(device >= 0 && device < ndev, MPQC_T_ERR_BAD_ARG, "device ordinal out of range");
  MPQC_T_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MPQC_T_CUDA(cudaGetDeviceProperties(&prop, device));
  double* out = nullptr;
  MPQC_T_CUDA(cudaMalloc(&out, 64));
  // which = 2: DMMA with ONE 8-warp block per SM (two warps per scheduler, the W-contraction kernel's occupancy)
  const int blocks = prop.multiProcessorCount * (which == 2 ? 1 : 4), threads = 256, iters = 20000;
  cudaEvent_t e0, e1;
  MPQC_T_CUDA(cudaEventCreate(&e0));
  MPQC_T_CUDA(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (which == 0 || which == 2) microbench_dmma_kernel<<<blocks, threads>>>(out, iters);
    else microbench_dfma_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    MPQC_T_CUDA(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  MPQC_T_CUDA(cudaGetLastError());
  double fl;
  if (which == 0 || which == 2) fl = (double)blocks * (threads / 32) * (double)iters * 8.0 * 512.0;
  else fl = (double)blocks * threads * (double)iters * 8.0 * 2.0;
  *tflops = fl / (best * 1e-3) * 1e-12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return MPQC_T_OK;
}

}  // extern "C"
