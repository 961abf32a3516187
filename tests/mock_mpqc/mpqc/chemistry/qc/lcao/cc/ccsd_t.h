// MOCK of the CCSD / CCSD_T / KeyVal / Energy surface the adapter uses (names and signatures as in the reference:
// ccsd.h:54-265,480-493; ccsd_t.h:33-197,2210-2244; keyval.h; properties/energy.h; exenv.h).  Not MPQC.
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <tiledarray.h>
namespace mpqc {
struct KeyVal { template <class T> T value(const std::string&, const T& def) const { return def; } };
struct ExEnv { static std::ostream& out0() { return std::cout; } };
struct Energy {};
struct Registry { void purge() {} };
struct LCAOFactory { Registry r; Registry& registry() { return r; } madness::World w; madness::World& world() { return w; } };
struct WfnWorld { madness::World w; madness::World& world() { return w; } };
namespace utility {
struct TRange1Engine {
  std::size_t get_occ() const { return 5; } std::size_t get_nfrozen() const { return 1; }
  std::size_t get_active_occ() const { return 4; } std::size_t get_vir() const { return 8; }
};
}  // namespace utility
namespace lcao {
template <class Tile, class Policy>
class CCSD {
 public:
  using TArray = TA::DistArray<Tile, Policy>;
  explicit CCSD(const KeyVal&) {}
  virtual ~CCSD() {}
  TArray t1() const { return TArray(); }
  TArray t2() const { return TArray(); }
  bool is_df() const { return true; }
  bool verbose() const { return false; }
  bool computed() const { return computed_; }
  std::shared_ptr<const Eigen::VectorXd> orbital_energy() const { return std::make_shared<Eigen::VectorXd>(); }
  std::shared_ptr<utility::TRange1Engine> trange1_engine() const { return std::make_shared<utility::TRange1Engine>(); }
  std::shared_ptr<WfnWorld> wfn_world() const { return std::make_shared<WfnWorld>(); }
  LCAOFactory& lcao_factory() { return f_; }
  struct Value { struct D { double operator[](int) const { return 0.0; } }; D derivs(int) const { return D(); } };
  Value get_value(Energy*) const { return Value(); }
  void set_value(Energy*, double) {}
  virtual void evaluate(Energy*) {}
 protected:
  const TArray get_Xab() { return TArray(); }
  const TArray get_Xij() { return TArray(); }
  const TArray get_Xai() { return TArray(); }
  bool computed_ = false;
  LCAOFactory f_;
};
template <class Tile, class Policy>
class CCSD_T : virtual public CCSD<Tile, Policy> {
 public:
  using TArray = TA::DistArray<Tile, Policy>;
  explicit CCSD_T(const KeyVal& kv) : CCSD<Tile, Policy>(kv) {}
 protected:
  void evaluate(Energy*) override {}
  const TArray get_aijk() { return TArray(); }
  const TArray get_abci() { return TArray(); }
  const TArray get_abij() { return TArray(); }
  double triples_energy_ = 0.0;
};
}  // namespace lcao
}  // namespace mpqc
