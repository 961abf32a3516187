// MOCK of cc/ccsd.h and of what it pulls in (KeyVal, Energy, ExEnv, OrbitalIndex, TRange1Engine, the factories): the
// base-class surface that the reference's REAL (patched) ccsd_t.h touches when CCSD_T's constructor and
// compute_ccsd_t_gpu() are instantiated.  Access specifiers and signatures restate the reference:
//   CCSD            ccsd.h:53-56 (bases), :86 (ctor), :120-142 (protected data, verbose_ :137), :144-148 (protected
//                   orbital_energy), :156-185 (public obsolete/t1/t2/is_df/verbose), :195-205 (protected evaluate),
//                   :479-493 (protected get_Xab/get_Xij/get_Xai)
//   Wavefunction    chemistry/qc/wfn/wfn.h:64-69 (wfn_world, obsolete, computed), :82 (protected computed_)
//   LCAOWavefunction lcao/wfn/lcao_wfn.h:76-82 (lcao_factory / ao_factory), :192-195 (trange1_engine)
//   TRange1Engine   lcao/expression/trange1_engine.h:60-72
// tests/test_integration_patch.py additionally checks, against the real ccsd.h, that every CCSD member the GPU code
// uses is declared public or protected there.  The CCSD_T side is NOT mocked: the real header is compiled.
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <tiledarray.h>
#include "mpqc/math/external/eigen/eigen.h"
#include "mpqc/util/core/exception.h"   // the reference's real header
#include "mpqc/util/misc/time.h"        // the reference's real header
namespace mpqc {
struct KeyVal {   // util/keyval/keyval.h: exists(), value<T>(key, default)
  bool exists(const std::string&) const { return false; }
  template <class T> T value(const std::string&, const T& def) const { return def; }
  template <class T> T value(const std::string&, const char* def) const { return T(def); }
};
struct ExEnv { static std::ostream& out0() { return std::cout; } };   // util/core/exenv.h
struct Energy {};                                                      // chemistry/qc/properties/energy.h
struct OrbitalIndex { explicit OrbitalIndex(const std::wstring&) {} }; // lcao/expression/orbital_index.h
struct Basis { std::size_t nfunctions() const { return 0; } };
struct BasisRegistry { std::shared_ptr<Basis> retrieve(const OrbitalIndex&) const { return std::make_shared<Basis>(); } };
struct Registry { void purge() {} };
struct WavefunctionWorld {
  madness::World w;
  madness::World& world() { return w; }
  std::shared_ptr<BasisRegistry> basis_registry() const { return std::make_shared<BasisRegistry>(); }
};
namespace utility {
struct TRange1Engine {   // lcao/expression/trange1_engine.h:60-72 (counts), :74-100 (tilings)
  std::size_t get_occ() const { return 5; }
  std::size_t get_nfrozen() const { return 1; }
  std::size_t get_active_occ() const { return 4; }
  std::size_t get_vir() const { return 8; }
  std::size_t get_all() const { return 13; }
  std::size_t get_occ_block_size() const { return 4; }
  std::size_t get_vir_block_size() const { return 8; }
  std::size_t get_active_occ_blocks() const { return 1; }
  std::size_t get_vir_blocks() const { return 1; }
  TA::TiledRange1 get_active_occ_tr1() const { return TA::TiledRange1(); }
  TA::TiledRange1 get_vir_tr1() const { return TA::TiledRange1(); }
};
TA::TiledRange1 compute_trange1(std::size_t, std::size_t);   // lcao/expression/trange1_engine.h
}  // namespace utility
namespace detail {
template <typename... Args> void parallel_print_range_info(Args&&...);   // util/misc/print.h family
}  // namespace detail
namespace util {
template <typename... Args> void print_progress(Args&&...);               // util/misc/print.h family
}  // namespace util
namespace math {
template <typename... Args> void create_diagonal_array_from_eigen(Args&&...);   // math/linalg/diagonal_array.h:16
}  // namespace math
namespace lcao {
template <class Array> class OrbitalSpace;   // lcao/expression/orbital_space.h (only named by the CPU reblock code)
template <class Tile, class Policy>
struct LCAOFactory {
  Registry r;
  madness::World w;
  Registry& registry() { return r; }
  madness::World& world() { return w; }
  TA::DistArray<Tile, Policy> compute(const std::wstring&) { return TA::DistArray<Tile, Policy>(); }
};
class Wavefunction {   // chemistry/qc/wfn/wfn.h
 public:
  virtual ~Wavefunction() {}
  const std::shared_ptr<WavefunctionWorld>& wfn_world() const { return wfn_world_; }
  virtual void obsolete() { computed_ = false; }
  bool computed() const { return computed_; }
 protected:
  bool computed_ = false;
 private:
  std::shared_ptr<WavefunctionWorld> wfn_world_ = std::make_shared<WavefunctionWorld>();
};
template <class Tile, class Policy>
class LCAOWavefunction : public Wavefunction {   // lcao/wfn/lcao_wfn.h
 public:
  explicit LCAOWavefunction(const KeyVal&) {}
  LCAOFactory<Tile, Policy>& lcao_factory() { return f_; }
  LCAOFactory<Tile, Policy>& ao_factory() { return f_; }
  const std::shared_ptr<const ::mpqc::utility::TRange1Engine>& trange1_engine() const { return tre_; }
 private:
  LCAOFactory<Tile, Policy> f_;
  std::shared_ptr<const ::mpqc::utility::TRange1Engine> tre_ = std::make_shared<::mpqc::utility::TRange1Engine>();
};
template <class Tile, class Policy>
class CCSD : public LCAOWavefunction<Tile, Policy> {
 public:
  using TArray = TA::DistArray<Tile, Policy>;
  CCSD() : LCAOWavefunction<Tile, Policy>(KeyVal()) {}
  CCSD(const KeyVal& kv) : LCAOWavefunction<Tile, Policy>(kv) {}
  virtual ~CCSD() {}
 protected:
  bool df_ = true;
  bool verbose_ = false;
  std::shared_ptr<const EigenVector<typename Tile::numeric_type>> f_pq_diagonal_;
 protected:
  std::shared_ptr<const EigenVector<typename Tile::numeric_type>> orbital_energy() { return f_pq_diagonal_; }
 public:
  void obsolete() override { LCAOWavefunction<Tile, Policy>::obsolete(); }
  TArray t1() const { return TArray(); }
  TArray t2() const { return TArray(); }
  bool is_df() const { return df_; }
  bool verbose() const { return verbose_; }
  struct Value { struct D { double operator[](int) const { return 0.0; } }; D derivs(int) const { return D(); } };
  Value get_value(Energy*) const { return Value(); }
  void set_value(Energy*, double) {}
 protected:
  virtual void evaluate(Energy*) {}
 protected:
  const TArray get_Xab() { return this->lcao_factory().compute(L"(Κ|G|a b)[inv_sqr]"); }
  const TArray get_Xij() { return this->lcao_factory().compute(L"(Κ|G|i j)[inv_sqr]"); }
  const TArray get_Xai() { return this->lcao_factory().compute(L"(Κ|G|a i)[inv_sqr]"); }
};
}  // namespace lcao
}  // namespace mpqc
