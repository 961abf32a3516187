// Type check of the reference's REAL ccsd_t.h after integration/mpqc_ccsd_t_gpu.patch, with ccsd_t_gpu_impl.h beside it
// (tests/test_integration_patch.py builds that tree in a temporary directory).  Instantiates, with the reference's own
// tile/policy pair (ccsd_t.cpp:12-13): the class (all member declarations) and compute_ccsd_t_gpu() (uses the class's
// private getters and the protected CCSD interface).
#include "mpqc/chemistry/qc/lcao/cc/ccsd_t.h"

typedef mpqc::lcao::CCSD_T<TA::TensorD, TA::SparsePolicy> T_sparse;
// (the constructor's patched keyword parsing is made of non-dependent expressions -- kv is a KeyVal -- so it is type
// checked when the template is parsed; instantiating the constructor would need the vtable and with it every CPU
// algorithm of the class, i.e. a full TiledArray expression engine)
template double T_sparse::compute_ccsd_t_gpu();
static_assert(sizeof(T_sparse) > 0, "class body instantiates");

// the second caller: CCSD_T_F12::evaluate (f12/ccsd_t_f12.h:47-69) calls the protected, NON-virtual
// CCSD_T::compute_ccsd_t() of its base and then reads triples_energy().  Both must stay reachable from a derived class
// (unevaluated operands: access and overload resolution are checked, the CPU algorithms' bodies are not instantiated).
#include <type_traits>
#include <utility>
struct SecondCaller : public T_sparse {
  explicit SecondCaller(const mpqc::KeyVal &kv);
  typedef decltype(std::declval<SecondCaller &>().T_sparse::compute_ccsd_t()) dispatch_result;
  typedef decltype(std::declval<SecondCaller &>().triples_energy()) energy_result;
  static_assert(std::is_same<dispatch_result, void>::value && std::is_same<energy_result, double>::value,
                "CCSD_T_F12 keeps compiling against the patched class");
};
