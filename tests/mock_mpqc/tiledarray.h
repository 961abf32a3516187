// MOCK of the small TiledArray surface integration/ccsd_t_gpu.h touches -- for a syntax/type check of the adapter
// only (tests/test_host_logic.py::test_adapter_header_compiles_against_mocks).  Not TiledArray.
#pragma once
#include <array>
#include <cstddef>
#include <memory>
#include <vector>
namespace madness {
struct Gop { void fence() {} template <class T> void sum(T&) {} };
struct World { Gop gop; int rank() const { return 0; } int size() const { return 1; } };
}  // namespace madness
namespace TA {
struct Range {
  std::vector<std::size_t> ext;
  std::vector<std::size_t> extent() const { return ext; }
  std::size_t volume() const { std::size_t n = 1; for (auto e : ext) n *= e; return n; }
  std::size_t rank() const { return ext.size(); }
  std::vector<std::vector<std::size_t>> coords;
  auto begin() const { return coords.begin(); }
  auto end() const { return coords.end(); }
};
struct TiledRange {
  Range er, tr;
  const Range& elements_range() const { return er; }
  const Range& tiles_range() const { return tr; }
};
struct TensorD {
  Range r; std::vector<double> d;
  const Range& range() const { return r; }
  const double* data() const { return d.data(); }
};
struct DensePolicy {};
struct SparsePolicy {};
template <class Tile, class Policy>
class DistArray {
 public:
  typedef Tile value_type;
  struct Future { Tile t; Tile get() const { return t; } };
  madness::World& world() { static madness::World w; return w; }
  void make_replicated() {}
  const TiledRange& trange() const { return tr_; }
  const Future* begin() const { return f_.data(); }
  const Future* end() const { return f_.data() + f_.size(); }
 private:
  TiledRange tr_; std::vector<Future> f_;
};
}  // namespace TA
namespace Eigen {
struct VectorXd { std::vector<double> v; const double* data() const { return v.data(); } };
}  // namespace Eigen
