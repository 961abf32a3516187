// libmpqc_t_cuda.so -- host driver + C ABI (include/mpqc_t.h) of the B200 (T) path.
//
// Replaces the body of CCSD_T::compute_ccsd_t() (ccsd_t.h:144-177) / compute_ccsd_t_coarse_grain
// (ccsd_t.h:200-711): integrals and amplitudes arrive as dense buffers, are re-laid-out once into
// occupied-major operand panels (relayout.cuh), and the (i>=j>=k) triple space is walked in batches
// of {W-contraction DMMA kernel (w_contract.cuh) -> fused energy kernel (t_energy.cuh)}.
// There is no CPU fallback: without a CUDA device every entry point returns MPQC_T_ERR_NO_DEVICE.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "driver.cuh"
#include "microbench.cuh"

// -------------------------------------------------------------------------------------------------
// C ABI (include/mpqc_t.h)
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

int64_t mpqc_t_shard_plan(int64_t o, int64_t first, int64_t stride, int64_t count, int32_t nranks, int32_t rank,
                          int32_t all_local, int32_t panel_block, int64_t* positions, int64_t capacity, int64_t* tail_begin) {
  if (o < 1 || o > 4096 || nranks < 1 || rank < 0 || rank >= nranks || first < 0 || panel_block < 0) {
    fail(MPQC_T_ERR_BAD_ARG, "bad shard plan arguments", __FILE__, __LINE__);
    return -1;
  }
  if (stride <= 0) stride = 1;
  const UnitIndex ux(o);
  const int64_t nt = ux.count();
  const int64_t avail = first < nt ? (nt - first + stride - 1) / stride : 0;
  if (count < 0 || count > avail) count = avail;
  std::vector<int64_t> mine;
  static_share(ux, first, stride, count, nranks, rank, all_local != 0, panel_block, mine);
  if (tail_begin) *tail_begin = static_share_end(count, nranks, all_local != 0, panel_block);
  if (positions)
    for (size_t q = 0; q < mine.size() && (int64_t)q < capacity; ++q) positions[q] = mine[q];
  return (int64_t)mine.size();
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

// One tiny all-reduce and all-gather per member right after the communicator exists: NCCL sets up its peer
// connections lazily on the first collective of each kind, which otherwise costs the first (T) call ~0.5 s.
static int comm_member_warmup(const CommMember& m, int nranks) {
  if (nranks <= 1 || !m.comm) return MPQC_T_OK;
  const NcclApi& nc = nccl_api();
  MPQC_T_CUDA(cudaSetDevice(m.device));
  cudaStream_t st = nullptr;
  MPQC_T_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = cuda_status(cudaMemsetAsync(m.scratch, 0, (size_t)(nranks + 1) * sizeof(double), st), "cudaMemsetAsync", __FILE__, __LINE__);
  if (rc == MPQC_T_OK)
    rc = nccl_status(nc.AllReduce(m.scratch, m.scratch, 1, kNcclFloat64, kNcclSum, m.comm, st), "ncclAllReduce(warm-up)", __FILE__, __LINE__);
  if (rc == MPQC_T_OK)
    rc = nccl_status(nc.AllGather(m.scratch + m.rank, m.scratch, 1, kNcclFloat64, m.comm, st), "ncclAllGather(warm-up)", __FILE__, __LINE__);
  if (rc == MPQC_T_OK) rc = cuda_status(cudaStreamSynchronize(st), "cudaStreamSynchronize(warm-up)", __FILE__, __LINE__);
  cudaStreamDestroy(st);
  return rc;
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
      if (rc == MPQC_T_OK) rc = comm_member_warmup(c->members[0], nranks);
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
      if (rc == MPQC_T_OK) {   // warm-up collectives, one thread per member (each blocks until all have joined)
        std::vector<std::thread> th;
        for (int g = 0; g < ngpu; ++g)
          th.emplace_back([&, g] {
            rcs[g] = comm_member_warmup(c->members[g], ngpu);
            if (rcs[g] != MPQC_T_OK) msgs[g] = last_error_string();
          });
        for (auto& t : th) t.join();
        for (int g = 0; g < ngpu && rc == MPQC_T_OK; ++g)
          if (rcs[g] != MPQC_T_OK) {
            last_error_string() = msgs[g];
            rc = rcs[g];
          }
      }
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
