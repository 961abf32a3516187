// MOCK of the TiledArray / MADNESS surface that the reference's ccsd_t.h and integration/ccsd_t_gpu_impl.h touch
// when CCSD_T's constructor and compute_ccsd_t_gpu() are instantiated.  For the type check of
// tests/test_integration_patch.py only.  Not TiledArray: every signature restates the public TiledArray / MADNESS API
// (tiledarray/dist_array.h, range.h, tiled_range.h; madness/world/worldgop.h) as the reference's own call sites use it
// (e.g. math/tensor/clr/cp_als.h:83-86, util/external/madworld/parallel_file.cpp:32-39, math/external/tiledarray/
// array_info.h:63).
#pragma once
#include <array>
#include <cstddef>
#include <memory>
#include <string>
#include <utility>
#include <vector>
namespace madness {
typedef int ProcessID;
struct WorldGopInterface {
  void fence() {}
  template <class T> void sum(T&) {}
  template <class T> void sum(T*, std::size_t) {}
  template <class T> void broadcast(T&, ProcessID) {}
  template <class T> void broadcast(T*, std::size_t, ProcessID) {}
};
}  // namespace madness
namespace SafeMPI {   // madness/world/safempi.h, as ccsd_t.h:290-292 uses it
struct Group { Group Incl(int, const int*) const { return Group(); } };
struct Intracomm {
  Group Get_group() const { return Group(); }
  Intracomm Create(const Group&) const { return Intracomm(); }
};
}  // namespace SafeMPI
namespace madness {
struct WorldMpiInterface { SafeMPI::Intracomm& comm() { static SafeMPI::Intracomm c; return c; } };
struct World {
  World() {}
  explicit World(const SafeMPI::Intracomm&) {}
  WorldGopInterface gop;
  WorldMpiInterface mpi;
  ProcessID rank() const { return 0; }
  ProcessID size() const { return 1; }
};
}  // namespace madness
namespace TiledArray {
struct Range {
  typedef std::vector<std::size_t> index_view;
  index_view lo, ext;
  const index_view& lobound() const { return lo; }
  const index_view& extent() const { return ext; }
  std::size_t volume() const { std::size_t n = 1; for (auto e : ext) n *= e; return n; }
  unsigned int rank() const { return (unsigned int)ext.size(); }
};
struct TiledRange1 {
  std::size_t n = 0;
  std::size_t extent() const { return n; }
  std::size_t tile_extent() const { return 1; }
  std::pair<std::size_t, std::size_t> tiles_range() const { return std::make_pair(std::size_t(0), n); }
};
struct TiledRange {
  Range er, tr;
  std::vector<TiledRange1> d;
  const Range& elements_range() const { return er; }
  const Range& tiles_range() const { return tr; }
  const TiledRange1& dim(std::size_t i) const { return d[i]; }
};
template <class T>
struct Tensor {
  typedef T value_type;
  typedef T numeric_type;
  Range r;
  std::vector<T> d;
  const Range& range() const { return r; }
  const T* data() const { return d.data(); }
};
typedef Tensor<double> TensorD;
struct DensePolicy {};
struct SparsePolicy {};
template <class Tile>
struct Future {
  Tile t;
  Tile get() const { return t; }
};
struct TsrExpr {   // stand-in for an annotated-array expression: result("a,b,i,j") = result("i,j,a,b")
  template <class E> TsrExpr& operator=(const E&) { return *this; }
};
template <class Tile, class Policy>
class DistArray {
 public:
  TsrExpr operator()(const std::string&) { return TsrExpr(); }
  TsrExpr operator()(const std::string&) const { return TsrExpr(); }
  typedef Tile value_type;
  typedef Policy policy_type;
  typedef const Future<Tile>* const_iterator;
  madness::World& world() const { static madness::World w; return w; }
  void make_replicated() {}
  const TiledRange& trange() const { return tr_; }
  const_iterator begin() const { return f_.data(); }
  const_iterator end() const { return f_.data() + f_.size(); }
  // mock only: lets a test build an array from explicit tiles (zero tiles of a sparse array are simply absent)
  void mock_set(const TiledRange& tr, const std::vector<Tile>& tiles) {
    tr_ = tr;
    f_.clear();
    for (const Tile& t : tiles) f_.push_back(Future<Tile>{t});
  }
 private:
  TiledRange tr_;
  std::vector<Future<Tile>> f_;
};
template <typename... Args> double dot(Args&&...);            // tiledarray/expressions
void set_default_world(madness::World&);                        // tiledarray/external/madness.h
madness::World& get_default_world();
}  // namespace TiledArray
namespace TA = TiledArray;
