// integration/ccsd_t_gpu_densify.h -- TiledArray DistArray -> one dense row-major host buffer, the only non-trivial host
// logic of the GPU (T) adapter (installed as src/mpqc/chemistry/qc/lcao/cc/ccsd_t_gpu_densify.h, included by
// ccsd_t_gpu_impl.h).  Depends on nothing but the TiledArray tile/range interface, so it is RUN (not only type-checked)
// on the CPU against a mock of that interface: tests/test_integration_patch.py::test_densify_scatters_tiles_correctly.
#ifndef MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_DENSIFY_H_
#define MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_DENSIFY_H_

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <thread>
#include <vector>

#include <tiledarray.h>

namespace mpqc {
namespace lcao {
namespace gpu_t {

/// copies one tile into its place in the dense row-major buffer: contiguous runs along the last dimension, offsets
/// advanced by an odometer over the leading dimensions (no per-element N-d index arithmetic)
template <typename Tile>
void scatter_tile(const Tile &tile, const std::vector<std::size_t> &stride, double *out) {
  const auto &range = tile.range();
  const std::size_t rank = range.rank();
  const auto lo = range.lobound();
  const auto ext = range.extent();
  const std::size_t run = ext[rank - 1];
  std::size_t base = 0, nrun = 1;
  for (std::size_t d = 0; d < rank; ++d) base += std::size_t(lo[d]) * stride[d];
  for (std::size_t d = 0; d + 1 < rank; ++d) nrun *= ext[d];
  std::vector<std::size_t> idx(rank, 0);
  const double *src = tile.data();
  std::size_t off = base;
  for (std::size_t r = 0; r < nrun; ++r) {
    std::memcpy(out + off, src, run * sizeof(double));
    src += run;
    // odometer over dimensions rank-2 .. 0
    for (std::size_t d = rank - 1; d-- > 0;) {
      off += stride[d];
      if (++idx[d] < std::size_t(ext[d])) break;
      off -= stride[d] * ext[d];
      idx[d] = 0;
    }
  }
}

/// gathers a (possibly sparse-policy, distributed) array into the dense row-major buffer `out` (volume doubles) on every
/// rank; zero tiles of a sparse-policy array stay zero (sparse_threshold 1e-20, mpqc_task.cpp:23-24).  Replication idiom
/// of math/tensor/clr/cp_als.h:83-84; tiles are scattered by a few host threads (disjoint destinations).
template <typename Tile, typename Policy>
void densify_into(TA::DistArray<Tile, Policy> array, double *out_data) {
  auto &world = array.world();
  array.make_replicated();
  world.gop.fence();
  const auto &trange = array.trange();
  const auto &erange = trange.elements_range();
  const std::size_t rank = erange.rank();
  const auto ext = erange.extent();
  std::vector<std::size_t> stride(rank, 1);
  for (std::size_t d = rank - 1; d > 0; --d) stride[d - 1] = stride[d] * std::size_t(ext[d]);
  std::memset(out_data, 0, std::size_t(erange.volume()) * sizeof(double));
  std::vector<Tile> tiles;
  for (auto it = array.begin(); it != array.end(); ++it) tiles.push_back(it->get());
  const std::size_t nthread = std::max<std::size_t>(1, std::min<std::size_t>(8, std::thread::hardware_concurrency()));
  std::vector<std::thread> pool;
  for (std::size_t t = 0; t < nthread; ++t)
    pool.emplace_back([&, t] {
      for (std::size_t i = t; i < tiles.size(); i += nthread) scatter_tile(tiles[i], stride, out_data);
    });
  for (auto &th : pool) th.join();
}

}  // namespace gpu_t
}  // namespace lcao
}  // namespace mpqc

#endif  // MPQC4_SRC_MPQC_CHEMISTRY_QC_CC_CCSD_T_GPU_DENSIFY_H_
