// Dense inputs -> occupied-major operand layouts: host tensors are staged through the handle's arena (each rank moves
// 1/N of a tensor over PCIe, one ncclAllGather over NVLink completes it), then re-laid-out by the kernels of relayout.cuh.
#pragma once

#include "comm.cuh"
#include "handle.cuh"
#include "relayout.cuh"

namespace {

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

// sum-all-reduce of n doubles held on the host through the member's pre-allocated device scratch (chunked), so the
// collective itself never allocates.  Result overwrites x on every rank.
int allreduce_host_vector(const CommView& cv, double* x, size_t n, cudaStream_t st, int op = kNcclSum) {
  const NcclApi& nc = nccl_api();
  for (size_t c0 = 0; c0 < n; c0 += kCommScratchDoubles) {
    const size_t cn = std::min(kCommScratchDoubles, n - c0);
    MPQC_T_CUDA(cudaMemcpyAsync(cv.scratch, x + c0, cn * sizeof(double), cudaMemcpyHostToDevice, st));
    MPQC_T_NCCL(nc.AllReduce(cv.scratch, cv.scratch, cn, kNcclFloat64, op, cv.comm, st));
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

// Every rank calls this at the same point, right after its allocations and before the first input collective: returns
// the local error if there is one, an error if ANY peer failed, MPQC_T_OK otherwise -- so a rank that ran out of memory
// inside an upload makes all ranks return instead of leaving its peers blocked in an all-gather.
int agree(const CommView& cv, int local_rc, cudaStream_t st, const char* what) {
  const int nfail = count_failed_ranks(cv, local_rc, st);
  if (local_rc != MPQC_T_OK) return local_rc;
  if (nfail != 0) return fail(nfail < 0 ? MPQC_T_ERR_NCCL : MPQC_T_ERR_INTERNAL, what, __FILE__, __LINE__);
  return MPQC_T_OK;
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
  // dense inputs: all o panels resident.  All allocations of this upload happen here, followed by an agreement among
  // the ranks, before the first input collective.
  const int rc_alloc = [&]() -> int {
    MPQC_T_TRY(alloc_operands(h, (int)o, arena_new));
    MPQC_T_TRY(arena_reserve(h, arena_need));
    return MPQC_T_OK;
  }();
  MPQC_T_TRY(agree(cv, rc_alloc, st, "another rank of the (T) communicator could not allocate its operand memory"));
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

}  // namespace
