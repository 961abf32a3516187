// RUNS integration/ccsd_t_gpu_densify.h (the adapter's tile scatter) on the TiledArray mock: arrays of rank 1..4 with
// ragged tilings and a missing (zero) tile, every element compared with a brute-force N-d index walk.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <tiledarray.h>

#include "mpqc/chemistry/qc/lcao/cc/ccsd_t_gpu_densify.h"

typedef TA::TensorD Tile;
typedef TA::DistArray<Tile, TA::SparsePolicy> Array;

static double value_at(const std::vector<std::size_t> &g) {
  double x = 0.5;
  for (std::size_t d = 0; d < g.size(); ++d) x = x * 37.0 + double(g[d] + 1) * (d + 1);
  return x;
}

// builds the array from per-dimension tile boundaries; tile number `skip` (in row-major tile order) is left out
static int check(const std::vector<std::vector<std::size_t>> &bounds, long skip) {
  const std::size_t rank = bounds.size();
  std::vector<std::size_t> ext(rank), ntile(rank);
  std::size_t volume = 1, ntiles = 1;
  for (std::size_t d = 0; d < rank; ++d) {
    ext[d] = bounds[d].back();
    ntile[d] = bounds[d].size() - 1;
    volume *= ext[d];
    ntiles *= ntile[d];
  }
  TA::TiledRange tr;
  tr.er.lo.assign(rank, 0);
  tr.er.ext = ext;
  tr.tr.lo.assign(rank, 0);
  tr.tr.ext = ntile;
  std::vector<Tile> tiles;
  std::vector<double> expect(volume, 0.0);
  for (std::size_t t = 0; t < ntiles; ++t) {
    std::vector<std::size_t> ti(rank);
    std::size_t rem = t;
    for (std::size_t d = rank; d-- > 0;) { ti[d] = rem % ntile[d]; rem /= ntile[d]; }
    Tile tile;
    tile.r.lo.resize(rank);
    tile.r.ext.resize(rank);
    std::size_t tv = 1;
    for (std::size_t d = 0; d < rank; ++d) {
      tile.r.lo[d] = bounds[d][ti[d]];
      tile.r.ext[d] = bounds[d][ti[d] + 1] - bounds[d][ti[d]];
      tv *= tile.r.ext[d];
    }
    tile.d.resize(tv);
    for (std::size_t e = 0; e < tv; ++e) {   // row-major inside the tile
      std::vector<std::size_t> g(rank);
      std::size_t r2 = e;
      for (std::size_t d = rank; d-- > 0;) { g[d] = tile.r.lo[d] + r2 % tile.r.ext[d]; r2 /= tile.r.ext[d]; }
      tile.d[e] = value_at(g);
      if ((long)t != skip) {
        std::size_t off = 0;
        for (std::size_t d = 0; d < rank; ++d) off = off * ext[d] + g[d];
        expect[off] = tile.d[e];
      }
    }
    if ((long)t != skip) tiles.push_back(tile);
  }
  Array a;
  a.mock_set(tr, tiles);
  std::vector<double> out(volume, -1.0);   // densify_into must overwrite everything, zero tiles included
  mpqc::lcao::gpu_t::densify_into(a, out.data());
  for (std::size_t i = 0; i < volume; ++i)
    if (out[i] != expect[i]) {
      std::printf("rank %zu: mismatch at %zu: %g vs %g\n", rank, i, out[i], expect[i]);
      return 1;
    }
  return 0;
}

int main() {
  int bad = 0;
  bad += check({{0, 3, 7}}, -1);                                               // rank 1
  bad += check({{0, 2, 5}, {0, 4, 6, 7}}, 3);                                  // rank 2, one zero tile
  bad += check({{0, 1, 4}, {0, 3}, {0, 2, 3, 8}}, -1);                         // rank 3
  bad += check({{0, 2, 5}, {0, 4}, {0, 1, 3}, {0, 4, 6}}, 5);                  // rank 4 (t2 / integral shape), zero tile
  bad += check({{0, 8, 16, 19}, {0, 8, 16, 19}, {0, 8, 16, 19}, {0, 4}}, 0);   // <ia|bc>-like: v = 19 in blocks of 8, o = 4
  std::printf(bad ? "densify: FAILED\n" : "densify: ok\n");
  return bad;
}
