// The triples loop of one device: batches of {W-contraction kernel -> fused energy kernel} over an explicit unit list;
// in panel-cache mode grouped by occupied-block triple with the needed operand panels built on the stream first.
#pragma once

#include "df_build.cuh"

namespace {

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

}  // namespace
