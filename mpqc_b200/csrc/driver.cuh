// One-shot driver shared by mpqc_t_energy[_df][_comm]: workers (one per local GPU), agreement collectives, input
// replication, unit sharding (static + work stealing, or by panel group), the final ncclAllReduce of the unit energies.
#pragma once

#include <atomic>
#include <functional>
#include <string>
#include <thread>

#include "run_units.cuh"

namespace {

int validate_problem(const mpqc_t_problem* p) {
  MPQC_T_CHECK(p != nullptr, MPQC_T_ERR_BAD_ARG, "problem is NULL");
  MPQC_T_CHECK(p->o >= 1 && p->v >= 1, MPQC_T_ERR_BAD_ARG, "o and v must be >= 1");
  MPQC_T_CHECK(p->o <= 4096 && p->v <= 2040, MPQC_T_ERR_BAD_ARG, "o <= 4096 and v <= 2040 supported");
  MPQC_T_CHECK(p->eps_occ && p->eps_vir && p->t1 && p->t2 && p->g_abij && p->g_aijk && p->g_abci,
               MPQC_T_ERR_BAD_ARG, "a tensor pointer is NULL");
  return MPQC_T_OK;
}

typedef std::function<int(mpqc_t_handle*, const CommView&, mpqc_t_stats*)> UploadFn;

// Static share of a job among W workers (host only; also exported as mpqc_t_shard_plan so the split can be tested
// without a GPU).  Job positions are 0-based indices into the job  first, first+stride, ...  (count of them).
//   panel_block == 0: unit-cyclic -- worker w takes positions w, w+W, ... below `static_n`; when all workers share a
//     process, static_n = the first 7/8 (a multiple of W) and the rest is handed out at run time by an atomic counter
//     (work stealing); otherwise static_n = count.
//   panel_block  > 0 (operand panel cache): shard by occupied-block triple, not by unit -- a worker that holds a
//     group's panels runs the whole group.  Groups are dealt largest-first to the least loaded worker: every rank
//     computes the same assignment.
int64_t static_share_end(int64_t count, int W, bool all_local, int panel_block) {
  if (panel_block > 0) return count;
  return (W > 1 && all_local) ? (count / 8) * 7 / W * W : count;
}

void static_share(const UnitIndex& ux, int64_t first, int64_t stride, int64_t count, int W, int wrank, bool all_local,
                  int panel_block, std::vector<int64_t>& mine) {
  mine.clear();
  if (panel_block <= 0) {
    const int64_t static_n = static_share_end(count, W, all_local, 0);
    for (int64_t q = wrank; q < static_n; q += W) mine.push_back(q);
    return;
  }
  std::vector<std::pair<int64_t, int64_t>> keyed((size_t)count);   // (group key, job position)
  for (int64_t q = 0; q < count; ++q) {
    int i, j, k;
    ux.triple(first + q * stride, i, j, k);
    keyed[(size_t)q] = std::make_pair(block_key(i, j, k, panel_block), q);
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
  std::vector<int64_t> load((size_t)W, 0);
  for (const auto& gsz : groups) {
    const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
    load[(size_t)w] += gsz.first;
    if (w == wrank)
      for (int64_t a = gsz.second; a < gsz.second + gsz.first; ++a) mine.push_back(keyed[(size_t)a].second);
  }
}

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
  std::atomic<int64_t> tail_next(static_share_end(count, W, all_local, 0));   // first position of the work-stealing tail
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
    if (rc == MPQC_T_OK) {
      // static share: unit-cyclic, or by occupied-block group in panel-cache mode (static_share above)
      const int pblock = h->panel_mode ? panel_block_edge(h) : 0;
      const int64_t my_static_n = static_share_end(count, W, all_local, pblock);   // == count in panel mode: no tail
      std::vector<int64_t> mine, idx;
      static_share(ux, opt.unit_first, stride, count, W, wrank, all_local, pblock, mine);
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
      while (rc == MPQC_T_OK && my_static_n < count) {
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
