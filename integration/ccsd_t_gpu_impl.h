// integration/ccsd_t_gpu_impl.h -- body of CCSD_T<Tile,Policy>::compute_ccsd_t_gpu(), the member function that
// integration/mpqc_ccsd_t_gpu.patch adds to the reference's own class
//   src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h
// (installed next to it as  src/mpqc/chemistry/qc/lcao/cc/ccsd_t_gpu_impl.h  and included at the end of the patched
// header).  It is an IN-CLASS change on purpose: the integral getters get_abij/get_aijk/get_abci (ccsd_t.h:2210-2244)
// and triples_energy_ (ccsd_t.h:72) are private members of CCSD_T, and CCSD_T_F12 (f12/ccsd_t_f12.h:59) calls the
// non-virtual, protected compute_ccsd_t() of its base -- a subclass could reach neither.  With approach == "gpu"
// dispatched inside compute_ccsd_t() (ccsd_t.h:157-171) both "CCSD(T)" and "CCSD(T)F12" get the GPU path.
//
// What this function does: gather the converged amplitudes and the three integral classes (or, for a density-fitted
// CCSD, the three-centre factors of ccsd.h:480-493) into dense page-locked host buffers, hand them through the C ABI
// of mpqc_t.h to libmpqc_t_cuda.so, and return the total E(T).  Every line of arithmetic lives behind the C ABI.
//
// Not compiled against real MPQC in this repository's container (TiledArray / MADNESS / Libint2 are absent there):
// tests/test_integration_patch.py applies the patch to the reference's real header, then type-checks the patched
// header + this file against mocks of TiledArray, MADNESS, Eigen and ccsd.h only (tests/mock_mpqc/).
#ifndef MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_IMPL_H_
#define MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_IMPL_H_

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "mpqc/chemistry/qc/lcao/cc/ccsd_t_gpu_densify.h"
#include "mpqc/util/core/exception.h"
#include "mpqc/util/misc/time.h"
#include "mpqc_t.h"  // include/mpqc_t.h of the mpqc_b200 repository

namespace mpqc {
namespace lcao {
namespace gpu_t {

/// status code of the C ABI -> the reference's exception classes (no exception crosses the ABI itself);
/// util/core/exception.h:135 (ProgrammingError), :190 (InputError), :277 (MemAllocFailed), :593 (FeatureDisabled)
inline void throw_on_error(int rc, const char *where) {
  if (rc == MPQC_T_OK) return;
  const std::string msg = std::string(where) + ": " + mpqc_t_strerror(rc) + "; " + mpqc_t_last_error();
  switch (rc) {
    case MPQC_T_ERR_BAD_ARG:
      throw InputError(msg.c_str(), __FILE__, __LINE__, "CCSD(T)");
    case MPQC_T_ERR_NO_DEVICE:
      throw FeatureDisabled(msg.c_str(), __FILE__, __LINE__, "GPU (T)");
    case MPQC_T_ERR_OOM:
      throw MemAllocFailed(msg.c_str(), __FILE__, __LINE__, 0);
    default:
      throw ProgrammingError(msg.c_str(), __FILE__, __LINE__);
  }
}

/// dense row-major host buffer in page-locked memory (full PCIe speed for the library's uploads)
class HostBuffer {
 public:
  HostBuffer() = default;
  explicit HostBuffer(std::size_t n) : n_(n) {
    void *p = nullptr;
    throw_on_error(mpqc_t_host_alloc(&p, std::max<std::size_t>(n, 1) * sizeof(double)), "mpqc_t_host_alloc");
    p_ = static_cast<double *>(p);
  }
  HostBuffer(HostBuffer &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; }
  HostBuffer &operator=(HostBuffer &&o) noexcept {
    std::swap(p_, o.p_);
    std::swap(n_, o.n_);
    return *this;
  }
  HostBuffer(const HostBuffer &) = delete;
  HostBuffer &operator=(const HostBuffer &) = delete;
  ~HostBuffer() { mpqc_t_host_free(p_); }
  double *data() { return p_; }
  const double *data() const { return p_; }
  std::size_t size() const { return n_; }

 private:
  double *p_ = nullptr;
  std::size_t n_ = 0;
};

/// gathers a (possibly sparse-policy, distributed) array into one dense row-major host buffer in page-locked memory on
/// every rank (ccsd_t_gpu_densify.h does the tile scatter)
template <typename Tile, typename Policy>
HostBuffer densify(TA::DistArray<Tile, Policy> array) {
  HostBuffer out(array.trange().elements_range().volume());
  densify_into(array, out.data());
  return out;
}

/// writes the dense problem in the MPQCT001 dump format (mpqc_b200/dump.py reads it): lets a real-molecule (T) be
/// replayed where MPQC/Libint are not installed.  Enabled by the keyword "gpu_dump_file".
inline void dump_problem(const std::string &path, const double *eps, std::size_t n_frozen, std::size_t o, std::size_t v,
                         const HostBuffer &t1, const HostBuffer &t2, const HostBuffer &g_abij, const HostBuffer &g_aijk,
                         const HostBuffer &g_abci) {
  std::FILE *f = std::fopen(path.c_str(), "wb");
  if (!f) throw FileOperationFailed("cannot open (T) dump file", __FILE__, __LINE__, path.c_str(), FileOperationFailed::OpenW);
  const int64_t hdr[4] = {int64_t(o), int64_t(v), int64_t(n_frozen), int64_t(n_frozen + o + v)};
  std::fwrite("MPQCT001", 1, 8, f);
  std::fwrite(hdr, sizeof(int64_t), 4, f);
  std::fwrite(eps, sizeof(double), n_frozen + o + v, f);
  for (const HostBuffer *a : {&t1, &t2, &g_abij, &g_aijk, &g_abci}) std::fwrite(a->data(), sizeof(double), a->size(), f);
  std::fclose(f);
}

/// the GPUs that share this (T): one MPI rank per GPU (rank mode; the NCCL id travels through world.gop.broadcast)
/// or, in a single-rank run, `ngpu` devices driven by this process (local mode).  Created once per wave function.
inline std::shared_ptr<mpqc_t_comm> make_comm(madness::World &world, int ngpu) {
  mpqc_t_comm *c = nullptr;
  if (world.size() > 1) {
    if (ngpu != 1)
      throw InputError("with more than one MPI rank the GPU (T) uses one GPU per rank", __FILE__, __LINE__, "ngpu");
    mpqc_t_unique_id id;
    std::memset(&id, 0, sizeof(id));
    if (world.rank() == 0) throw_on_error(mpqc_t_comm_unique_id(&id), "mpqc_t_comm_unique_id");
    world.gop.broadcast(id.internal, sizeof(id.internal), 0);
    const int ndev = std::max(1, mpqc_t_device_count());
    throw_on_error(mpqc_t_comm_create_rank(&c, world.size(), world.rank(), &id, world.rank() % ndev),
                   "mpqc_t_comm_create_rank");
  } else {
    throw_on_error(mpqc_t_comm_create_local(&c, ngpu, nullptr), "mpqc_t_comm_create_local");
  }
  return std::shared_ptr<mpqc_t_comm>(c, [](mpqc_t_comm *p) { mpqc_t_comm_destroy(p); });
}

}  // namespace gpu_t

/// approach == "gpu": replaces compute_ccsd_t_coarse_grain / _fine_grain / _straight (ccsd_t.h:200-1170) by one call
/// into libmpqc_t_cuda.so.  Collective over the MADWorld like the functions it replaces; returns the TOTAL E(T) on
/// every rank (the library's ncclAllReduce over NVLink replaces global_world.gop.sum, ccsd_t.h:692).
template <typename Tile, typename Policy>
double CCSD_T<Tile, Policy>::compute_ccsd_t_gpu() {
  auto &world = this->wfn_world()->world();
  auto tre = this->trange1_engine();
  const std::size_t n_occ = tre->get_occ(), n_frozen = tre->get_nfrozen();
  const std::size_t o = tre->get_active_occ(), v = tre->get_vir();
  const auto eps = this->orbital_energy();  // diagonal of the Fock matrix over ALL MOs, frozen core first (ccsd.h:141-148)
  const double *eps_occ = eps->data() + n_frozen;  // eps[i + n_frozen]   ccsd_t.h:2306-2311
  const double *eps_vir = eps->data() + n_occ;     // eps[a + n_occ]

  if (!gpu_comm_) gpu_comm_ = gpu_t::make_comm(world, gpu_ngpu_);

  mpqc_t_options opt;
  std::memset(&opt, 0, sizeof(opt));
  opt.unit_count = -1;  // the whole (i >= j >= k) list, sharded over the communicator (replaces ccsd_t.h:477-480)
  opt.batch = gpu_batch_;

  gpu_t::HostBuffer t1 = gpu_t::densify(this->t1());  // [a][i]        ccsd.h:165-171
  gpu_t::HostBuffer t2 = gpu_t::densify(this->t2());  // [a][b][i][j]  ccsd.h:173-179
  double e_t = 0.0;
  mpqc_t_stats st;
  std::memset(&st, 0, sizeof(st));
  if (gpu_df_ && this->is_df()) {
    // density-fitted hand-off: the v^3 o tensor is never formed on the host; the library assembles the three integral
    // classes on the device from the three-centre factors the CCSD already holds (ccsd.h:480-493)
    TArray Xab = this->get_Xab(), Xij = this->get_Xij(), Xai = this->get_Xai();
    const std::size_t naux = Xab.trange().dim(0).extent();
    gpu_t::HostBuffer xab = gpu_t::densify(Xab);  // [K][a][b]
    gpu_t::HostBuffer xij = gpu_t::densify(Xij);  // [K][i][j]
    gpu_t::HostBuffer xai = gpu_t::densify(Xai);  // [K][a][i]
    mpqc_t_df_problem p = {int64_t(o), int64_t(v), int64_t(naux), eps_occ,    eps_vir,
                           t1.data(),  t2.data(),  xab.data(),    xij.data(), xai.data()};
    gpu_t::throw_on_error(mpqc_t_energy_df_comm(gpu_comm_.get(), &p, &opt, &e_t, &st), "mpqc_t_energy_df_comm");
  } else {
    // dense blocks in exactly the layouts of ccsd_t.h:2219,2233,2242 (no reblock(): CPU tiling is irrelevant here)
    gpu_t::HostBuffer g_abij = gpu_t::densify(get_abij());
    gpu_t::HostBuffer g_aijk = gpu_t::densify(get_aijk());
    gpu_t::HostBuffer g_abci = gpu_t::densify(get_abci());
    if (!gpu_dump_file_.empty() && world.rank() == 0)
      gpu_t::dump_problem(gpu_dump_file_, eps->data(), n_frozen, o, v, t1, t2, g_abij, g_aijk, g_abci);
    mpqc_t_problem p = {int64_t(o), int64_t(v),  eps_occ,       eps_vir,      t1.data(),
                        t2.data(),  g_abij.data(), g_aijk.data(), g_abci.data()};
    gpu_t::throw_on_error(mpqc_t_energy_comm(gpu_comm_.get(), &p, &opt, &e_t, &st), "mpqc_t_energy_comm");
  }
  if (this->verbose_ && world.rank() == 0)
    ExEnv::out0() << "(T) GPU: upload " << st.seconds_upload << " S, relayout " << st.seconds_relayout << " S, triples "
                  << st.seconds_compute << " S, " << st.units << " units on this rank" << std::endl;
  return e_t;
}

}  // namespace lcao
}  // namespace mpqc

#endif  // MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_IMPL_H_
