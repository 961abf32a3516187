// Density-fitted inputs: factor GEMMs on the W-contraction kernel's plain NT-GEMM mode, operand panel pool and LRU cache.
#pragma once

#include "upload.cuh"

namespace {

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
  if (cv.nranks > 1) {
    // the choice depends on this device's free memory, and it decides how the job is sharded (by unit or by panel
    // group): all ranks must take the SAME one -- the smallest pool any rank asks for
    double np = (double)npanel;
    MPQC_T_TRY(allreduce_host_vector(cv, &np, 1, st, kNcclMin));
    npanel = (int)(np + 0.5);
  }
  const bool panel_mode = npanel < o;
  // all allocations of this upload, then an agreement among the ranks, before the first input collective
  DevBuf xijt;
  const size_t arena_need = staged_size((size_t)v * o, on_device, 1) + staged_size((size_t)v * v * o * o, on_device, cv.nranks) +
                            staged_size((size_t)naux * v * v, on_device, cv.nranks) +
                            staged_size((size_t)naux * o * o, on_device, 1) + staged_size((size_t)naux * v * o, on_device, cv.nranks);
  const int rc_alloc = [&]() -> int {
    const double arena_new = arena_need > h->arena_cap ? (double)arena_need * 8.0 : 0.0;
    MPQC_T_TRY(alloc_operands(h, npanel, factors + arena_new + (panel_mode ? t2_bytes : 0.0)));
    MPQC_T_TRY(arena_reserve(h, arena_need));
    MPQC_T_CUDA(cudaMalloc(&h->XaiT, (size_t)o * v * Kx * sizeof(double)));
    MPQC_T_CUDA(cudaMalloc(&h->XabT, (size_t)v * v * Kx * sizeof(double)));
    MPQC_T_TRY(xijt.alloc((size_t)o * o * Kx));
    if (panel_mode) {
      MPQC_T_CUDA(cudaMalloc(&h->T2raw, (size_t)v * v * o * o * sizeof(double)));
      MPQC_T_CUDA(cudaMalloc(&h->slot_map_dev, (size_t)o * sizeof(int)));
    }
    return MPQC_T_OK;
  }();
  MPQC_T_TRY(agree(cv, rc_alloc, st, "another rank of the (T) communicator could not allocate its operand memory"));
  MPQC_T_CUDA(cudaMemsetAsync(h->A, 0, (size_t)npanel * v * v * Kp * sizeof(double), st));
  if (h->flat) MPQC_T_CUDA(cudaMemsetAsync(h->AT, 0, (size_t)npanel * v * v * Kp * sizeof(double), st));
  MPQC_T_CUDA(cudaMemsetAsync(h->B, 0, (size_t)o * o * v * Kp * sizeof(double), st));

  const double tc = now_s();
  const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  MPQC_T_CUDA(cudaMemcpyAsync(h->eps_occ, p->eps_occ, o * sizeof(double), kind, st));
  MPQC_T_CUDA(cudaMemcpyAsync(h->eps_vir, p->eps_vir, v * sizeof(double), kind, st));
  if (!on_device) h2d += (o + v) * 8;
  {
    Staged t1, t2, xab, xij, xai;
    CommView solo;
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
      MPQC_T_CUDA(cudaMemcpyAsync(h->T2raw, t2.ptr, (size_t)v * v * o * o * sizeof(double), cudaMemcpyDeviceToDevice, st));
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
  auto evictable = [&](int q) {
    const int xq = h->x_of_slot[(size_t)q];
    return xq < 0 || !wanted[(size_t)xq];
  };
  auto assign = [&](int x, int slot) {
    if (h->x_of_slot[(size_t)slot] >= 0) h->slot_of[(size_t)h->x_of_slot[(size_t)slot]] = -1;
    h->x_of_slot[(size_t)slot] = x;
    h->slot_of[(size_t)x] = slot;
    h->slot_stamp[(size_t)slot] = h->stamp;
    changed = true;
  };
  // missing panels, as runs of consecutive occupied indices (a group's missing panels are usually one whole block)
  for (size_t a = 0; a < need.size();) {
    if (h->slot_of[(size_t)need[a]] >= 0) {
      h->slot_stamp[(size_t)h->slot_of[(size_t)need[a]]] = h->stamp;
      ++a;
      continue;
    }
    size_t b = a + 1;
    while (b < need.size() && need[b] == need[b - 1] + 1 && h->slot_of[(size_t)need[b]] < 0) ++b;
    const int len = (int)(b - a);
    // a window of `len` consecutive evictable slots lets the run be built by ONE flattened-row GEMM (build_panels);
    // among the candidates take the one whose most recently used slot is oldest (free slots count as oldest)
    int best = -1;
    int64_t best_age = 0;
    for (int s0 = 0; len > 1 && s0 + len <= h->npanel; ++s0) {
      int64_t newest = -1;
      bool ok = true;
      for (int t = 0; t < len && ok; ++t) {
        ok = evictable(s0 + t);
        if (ok && h->x_of_slot[(size_t)(s0 + t)] >= 0) newest = std::max(newest, h->slot_stamp[(size_t)(s0 + t)]);
      }
      if (ok && (best < 0 || newest < best_age)) {
        best = s0;
        best_age = newest;
      }
    }
    if (best >= 0) {
      for (int t = 0; t < len; ++t) assign(need[a + (size_t)t], best + t);
      MPQC_T_TRY(build_panels(h, need[a], len, best, launches));
    } else {
      // one panel at a time: a free slot, else the least recently used slot whose panel this group does not need
      for (size_t c = a; c < b; ++c) {
        int victim = -1;
        for (int q = 0; q < h->npanel; ++q) {
          if (!evictable(q)) continue;
          if (h->x_of_slot[(size_t)q] < 0) { victim = q; break; }
          if (victim < 0 || h->slot_stamp[(size_t)q] < h->slot_stamp[(size_t)victim]) victim = q;
        }
        MPQC_T_CHECK(victim >= 0, MPQC_T_ERR_INTERNAL, "no evictable panel slot");
        assign(need[c], victim);
        MPQC_T_TRY(build_panels(h, need[c], 1, victim, launches));
      }
    }
    a = b;
  }
  if (changed)   // pageable source: the runtime stages it before returning, so slot_of may change again right away
    MPQC_T_CUDA(cudaMemcpyAsync(h->slot_map_dev, h->slot_of.data(), (size_t)h->o * sizeof(int), cudaMemcpyHostToDevice,
                                h->stream));
  return MPQC_T_OK;
}

}  // namespace
