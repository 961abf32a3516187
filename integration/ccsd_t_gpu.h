// integration/ccsd_t_gpu.h -- MPQC-side adapter: the reference's CCSD(T) wavefunction with the body of
// compute_ccsd_t() replaced by one call into libmpqc_t_cuda.so (include/mpqc_t.h).
//
// NOT compiled in this repository's container (TiledArray / MADNESS / Libint2 are absent there); it is the
// binding a maintainer adds to an MPQC build tree.  Drop it next to
//   src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h
// and register it INSTEAD of the stock class (the registry asserts one entry per key,
// util/keyval/keyval.h:135) -- see INTEGRATION.md for the three-line change to ccsd_t.cpp / CMakeLists.txt.
//
// What stays from the reference: class hierarchy (CCSD_T : CCSD : LCAOWavefunction, Provides<Energy>), the
// KeyVal constructor and every keyword (ccsd_t.h:89-132), evaluate() (ccsd_t.h:179-197), the integral getters
// (ccsd_t.h:2210-2244), the log lines (ccsd_t.h:175,194).  What changes: compute_ccsd_t() (ccsd_t.h:144-177).
#ifndef MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_H_
#define MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_H_

#include <tiledarray.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "mpqc/chemistry/qc/lcao/cc/ccsd_t.h"
#include "mpqc/util/core/exception.h"
#include "mpqc/util/misc/time.h"
#include "mpqc_t.h"  // this repository's include/mpqc_t.h

namespace mpqc {
namespace lcao {

template <typename Tile, typename Policy>
class CCSD_T_GPU : public CCSD_T<Tile, Policy> {
 public:
  using TArray = TA::DistArray<Tile, Policy>;

  /// same keywords as CCSD_T plus: "ngpu" (devices per MPI rank, default 1), "gpu_batch" (0 = auto),
  /// "gpu_dump_file" (write the dense (T) inputs for replay).  With a density-fitted CCSD (is_df()) the three
  /// Xab/Xij/Xai factors (ccsd.h:480-493) can be passed through mpqc_t_energy_df instead of the dense
  /// integrals (same call shape) so the v^3 o tensor is never gathered on the host.
  /// "approach" coarse|fine|straight all map to the GPU path (one exact sum); "laplace" stays on the CPU.
  explicit CCSD_T_GPU(const KeyVal &kv)
      : CCSD<Tile, Policy>(kv), CCSD_T<Tile, Policy>(kv),
        ngpu_(kv.value<int>("ngpu", 1)), gpu_batch_(kv.value<int>("gpu_batch", 0)),
        laplace_(kv.value<std::string>("approach", "coarse") == "laplace"),
        dump_file_(kv.value<std::string>("gpu_dump_file", "")) {}

 protected:
  /// gathers a (possibly sparse-policy, distributed) array into one dense row-major host buffer; zero tiles
  /// stay zero (sparse_threshold 1e-20).  Idiom of math/tensor/clr/cp_als.h:83-86.
  static std::vector<double> densify(TArray array) {
    auto &world = array.world();
    array.make_replicated();
    world.gop.fence();
    const auto &trange = array.trange();
    const auto extent = trange.elements_range().extent();
    std::size_t n = trange.elements_range().volume();
    std::vector<double> out(n, 0.0);
    const std::size_t rank = trange.tiles_range().rank();
    std::vector<std::size_t> stride(rank, 1);
    for (std::size_t d = rank - 1; d > 0; --d) stride[d - 1] = stride[d] * extent[d];
    for (auto it = array.begin(); it != array.end(); ++it) {
      const Tile tile = it->get();
      const auto &r = tile.range();
      std::size_t idx = 0;
      for (const auto &coord : r) {
        std::size_t off = 0;
        for (std::size_t d = 0; d < rank; ++d) off += coord[d] * stride[d];
        out[off] = tile.data()[idx++];
      }
    }
    return out;
  }

  /// writes the dense problem in the MPQCT001 dump format (mpqc_b200/dump.py reads it): lets a real-molecule (T)
  /// be replayed where MPQC/Libint are not installed.  Enabled by the keyword "gpu_dump_file".
  static void dump_t_problem(const std::string &path, const Eigen::VectorXd &eps, std::size_t n_frozen, std::size_t o,
                             std::size_t v, const std::vector<double> &t1, const std::vector<double> &t2,
                             const std::vector<double> &g_abij, const std::vector<double> &g_aijk,
                             const std::vector<double> &g_abci) {
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) throw FileOperationFailed("cannot open (T) dump file", __FILE__, __LINE__, path.c_str(),
                                      FileOperationFailed::OpenW);
    const int64_t hdr[4] = {int64_t(o), int64_t(v), int64_t(n_frozen), int64_t(n_frozen + o + v)};
    std::fwrite("MPQCT001", 1, 8, f);
    std::fwrite(hdr, sizeof(int64_t), 4, f);
    std::fwrite(eps.data(), sizeof(double), n_frozen + o + v, f);
    for (const auto *a : {&t1, &t2, &g_abij, &g_aijk, &g_abci}) std::fwrite(a->data(), sizeof(double), a->size(), f);
    std::fclose(f);
  }

  /// replaces CCSD_T::compute_ccsd_t (ccsd_t.h:144-177)
  void compute_ccsd_t_gpu() {
    auto &world = this->wfn_world()->world();
    auto time0 = mpqc::fenced_now(world);
    this->lcao_factory().registry().purge();  // ccsd_t.h:149
    ExEnv::out0() << "\nBegining CCSD(T) " << std::endl;

    auto tre = this->trange1_engine();
    const std::size_t n_occ = tre->get_occ(), n_frozen = tre->get_nfrozen();
    const std::size_t o = tre->get_active_occ(), v = tre->get_vir();

    // dense blocks in the layouts of ccsd_t.h:2219,2233,2242 (no reblock(): tiling is irrelevant on the GPU)
    std::vector<double> t1 = densify(this->t1());
    std::vector<double> t2 = densify(this->t2());
    std::vector<double> g_abij = densify(this->get_abij());
    std::vector<double> g_aijk = densify(this->get_aijk());
    std::vector<double> g_abci = densify(this->get_abci());
    const Eigen::VectorXd &eps = *this->orbital_energy();  // all MOs incl. frozen core (ccsd.h:141-148)

    if (!dump_file_.empty() && world.rank() == 0)
      dump_t_problem(dump_file_, eps, n_frozen, o, v, t1, t2, g_abij, g_aijk, g_abci);

    mpqc_t_problem p;
    p.o = static_cast<int64_t>(o);
    p.v = static_cast<int64_t>(v);
    p.eps_occ = eps.data() + n_frozen;  // eps[i + n_frozen]   ccsd_t.h:2306-2311
    p.eps_vir = eps.data() + n_occ;     // eps[a + n_occ]
    p.t1 = t1.data();
    p.t2 = t2.data();
    p.g_abij = g_abij.data();
    p.g_aijk = g_aijk.data();
    p.g_abci = g_abci.data();

    mpqc_t_options opt = {};
    opt.ngpu = ngpu_;
    opt.batch = gpu_batch_;
    // one MPI rank per GPU (or per group of ngpu GPUs): shard the (i>=j>=k) units like ccsd_t.h:477-480
    opt.unit_first = world.rank();
    opt.unit_stride = world.size();
    opt.unit_count = -1;
    std::vector<int32_t> devs(ngpu_);
    for (int g = 0; g < ngpu_; ++g) devs[g] = (world.rank() * ngpu_ + g) % std::max(1, mpqc_t_device_count());
    opt.device_ids = devs.data();

    double e_partial = 0.0;
    mpqc_t_stats st;
    const int rc = mpqc_t_energy(&p, &opt, &e_partial, &st);
    switch (rc) {  // no exception crosses the C ABI; map status codes here
      case MPQC_T_OK: break;
      case MPQC_T_ERR_BAD_ARG: throw InputError(mpqc_t_last_error(), __FILE__, __LINE__, "CCSD(T)");
      case MPQC_T_ERR_NO_DEVICE: throw FeatureDisabled(mpqc_t_last_error(), __FILE__, __LINE__);
      case MPQC_T_ERR_OOM: throw MemAllocFailed(mpqc_t_last_error(), __FILE__, __LINE__, 0);
      default: throw ProgrammingError(mpqc_t_last_error(), __FILE__, __LINE__);
    }
    world.gop.sum(e_partial);  // the path's one collective, ccsd_t.h:692
    this->triples_energy_ = e_partial;

    auto time1 = mpqc::fenced_now(world);
    ExEnv::out0() << "(T) Energy: " << this->triples_energy_ << " Time: " << mpqc::duration_in_s(time0, time1)
                  << " S \n";  // ccsd_t.h:175
    if (this->verbose())
      ExEnv::out0() << "(T) GPU: upload " << st.seconds_upload << " S, relayout " << st.seconds_relayout
                    << " S, triples " << st.seconds_compute << " S, " << st.flops / st.seconds_compute * 1e-12
                    << " TFLOP/s\n";
  }

  /// ccsd_t.h:179-197 with the one call swapped
  void evaluate(Energy *result) override {
    if (laplace_) return CCSD_T<Tile, Policy>::evaluate(result);  // approximate method: reference CPU code
    if (!this->computed()) {
      auto &world = this->lcao_factory().world();
      CCSD<Tile, Policy>::evaluate(result);
      double ccsd_energy = this->get_value(result).derivs(0)[0];
      auto time0 = mpqc::fenced_now(world);
      compute_ccsd_t_gpu();
      this->computed_ = true;
      this->set_value(result, ccsd_energy + this->triples_energy_);
      auto time1 = mpqc::fenced_now(world);
      ExEnv::out0() << "(T) Time in CCSD(T): " << mpqc::duration_in_s(time0, time1) << " S" << std::endl;
    }
  }

 private:
  int ngpu_;
  int gpu_batch_;
  bool laplace_;
  std::string dump_file_;
};

}  // namespace lcao
}  // namespace mpqc

#endif  // MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_H_
