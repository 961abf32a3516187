/*
 * Plain-C restatement of the reference's (T) algorithm.  TEST INFRASTRUCTURE ONLY (see the header
 * of oracle/ccsd_t_oracle.py for who may use it and for the parity-pinning status).
 *
 * Follows, loop for loop:
 *   compute_ccsd_t_straight      /root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:1127-1170
 *   CCSD_T_Reduce::operator()    ccsd_t.h:2286-2334   (row-major walk a,b,c,i,j,k; k fastest)
 *   CCSD_T_ReduceSymm::operator()ccsd_t.h:2350-2431   (c<=b<=a, weights 2 / 1 / 0)
 * The TiledArray contraction of :1142-1143 is restated as explicit sums; no BLAS.
 * O(o^3 v^3) memory, O(o^3 v^3 (v+o)) time: small cases only.
 *
 * Layouts (row-major): t1[v][o], t2[v][v][o][o], g_abij[v][v][o][o], g_aijk[v][o][o][o],
 * g_abci[v][v][v][o]; eps = full orbital-energy vector (frozen, active occ, virt), as the reducers
 * index it: eps[i + n_frozen], eps[a + n_occ].
 */
#include <stdint.h>
#include <stdlib.h>

#define IDX6(a, b, c, i, j, k) ((((((int64_t)(a) * v + (b)) * v + (c)) * o + (i)) * o + (j)) * o + (k))

/* t3(a,b,c,i,j,k) = sum_d g_dabi(d,a,b,i) t2(d,c,j,k) - sum_l g_cjkl(c,j,k,l) t2(a,b,i,l)   :1142-1143 */
static void build_x(int o, int v, const double* t2, const double* g_aijk, const double* g_abci, double* x) {
  for (int a = 0; a < v; ++a)
    for (int b = 0; b < v; ++b)
      for (int c = 0; c < v; ++c)
        for (int i = 0; i < o; ++i)
          for (int j = 0; j < o; ++j)
            for (int k = 0; k < o; ++k) {
              double s = 0.0;
              for (int d = 0; d < v; ++d)
                s += g_abci[(((int64_t)d * v + a) * v + b) * o + i] * t2[(((int64_t)d * v + c) * o + j) * o + k];
              for (int l = 0; l < o; ++l)
                s -= g_aijk[(((int64_t)c * o + j) * o + k) * o + l] * t2[(((int64_t)a * v + b) * o + i) * o + l];
              x[IDX6(a, b, c, i, j, k)] = s;
            }
}

/* Denominator-weighted plain sum: what CCSD_T_Reduce::operator() (ccsd_t.h:2286-2334) computes for one tile covering
 * all indices -- sum over the row-major (a,b,c,i,j,k) walk of tile / (eps_i + eps_j + eps_k - eps_a - eps_b - eps_c),
 * occupied energies taken at [n_frozen + i], virtual ones at [n_occ + a]. */
double mpqc_oracle_reduce(int o, int v, int n_occ, int n_frozen, const double* eps, const double* tile) {
  const double* eo = eps + n_frozen;
  const double* ev = eps + n_occ;
  double total = 0.0;
  int64_t pos = 0;
  for (int a = 0; a < v; ++a)
    for (int b = 0; b < v; ++b)
      for (int c = 0; c < v; ++c) {
        const double virt = ev[a] + ev[b] + ev[c];
        for (int i = 0; i < o; ++i)
          for (int j = 0; j < o; ++j)
            for (int k = 0; k < o; ++k, ++pos) {
              const double denom = (eo[i] - virt) + eo[j] + eo[k];   /* same association order as :2311-2320 */
              total += (1.0 / denom) * tile[pos];
            }
      }
  return total;
}

/* Symmetry-restricted sum: what CCSD_T_ReduceSymm::operator() (ccsd_t.h:2350-2431) computes for one tile covering all
 * indices -- only c <= b <= a contribute, with weight 2 when a, b, c are all different, 0 when all equal, 1 otherwise
 * (the if/else ladder of :2410-2421). */
double mpqc_oracle_reduce_symm(int o, int v, int n_occ, int n_frozen, const double* eps, const double* tile) {
  const double* eo = eps + n_frozen;
  const double* ev = eps + n_occ;
  double total = 0.0;
  for (int a = 0; a < v; ++a)
    for (int b = 0; b <= a; ++b)
      for (int c = 0; c <= b; ++c) {
        double weight = 1.0;
        if (a != b && b != c && a != c) weight = 2.0;
        else if (a == b && b == c) weight = 0.0;
        const double virt = ev[a] + ev[b] + ev[c];
        for (int i = 0; i < o; ++i)
          for (int j = 0; j < o; ++j)
            for (int k = 0; k < o; ++k) {
              const double denom = (eo[i] - virt) + eo[j] + eo[k];
              total += weight * ((1.0 / denom) * tile[IDX6(a, b, c, i, j, k)]);
            }
      }
  return total;
}

/* mode 0: straight (full reduce / 3, :1163-1168); mode 1: same tensors through ReduceSymm (the
 * boundary-block branch of the coarse loop, :629-638, with one block covering all virtuals). */
double mpqc_oracle_straight(int o, int v, int n_frozen, const double* eps, const double* t1, const double* t2,
                            const double* g_abij, const double* g_aijk, const double* g_abci, int mode) {
  const int n_occ = n_frozen + o;
  const int64_t n6 = (int64_t)v * v * v * o * o * o;
  double* x = (double*)malloc(sizeof(double) * n6);
  double* t3 = (double*)malloc(sizeof(double) * n6);
  double* res = (double*)malloc(sizeof(double) * n6);
  if (!x || !t3 || !res) { free(x); free(t3); free(res); return 0.0 / 0.0; }
  build_x(o, v, t2, g_aijk, g_abci, x);
  /* :1144-1146 */
  for (int a = 0; a < v; ++a) for (int b = 0; b < v; ++b) for (int c = 0; c < v; ++c)
    for (int i = 0; i < o; ++i) for (int j = 0; j < o; ++j) for (int k = 0; k < o; ++k)
      t3[IDX6(a, b, c, i, j, k)] = x[IDX6(a, b, c, i, j, k)] + x[IDX6(a, c, b, i, k, j)] + x[IDX6(c, a, b, k, i, j)] +
                                   x[IDX6(c, b, a, k, j, i)] + x[IDX6(b, c, a, j, k, i)] + x[IDX6(b, a, c, j, i, k)];
  for (int a = 0; a < v; ++a) for (int b = 0; b < v; ++b) for (int c = 0; c < v; ++c)
    for (int i = 0; i < o; ++i) for (int j = 0; j < o; ++j) for (int k = 0; k < o; ++k) {
      /* v3, :1150-1152: y(abcijk) = g_abij(a,b,i,j) t1(c,k);  v3 = y + y(bcajki) + y(acbikj) */
      const double v3 = g_abij[(((int64_t)a * v + b) * o + i) * o + j] * t1[c * o + k] +
                        g_abij[(((int64_t)b * v + c) * o + j) * o + k] * t1[a * o + i] +
                        g_abij[(((int64_t)a * v + c) * o + i) * o + k] * t1[b * o + j];
      const double w = t3[IDX6(a, b, c, i, j, k)];
      /* :1163-1167 */
      const double z = 4.0 * w + t3[IDX6(a, b, c, k, i, j)] + t3[IDX6(a, b, c, j, k, i)] -
                       2.0 * (t3[IDX6(a, b, c, k, j, i)] + t3[IDX6(a, b, c, i, k, j)] + t3[IDX6(a, b, c, j, i, k)]);
      res[IDX6(a, b, c, i, j, k)] = (w + v3) * z;
    }
  double e;
  if (mode == 0) e = mpqc_oracle_reduce(o, v, n_occ, n_frozen, eps, res) / 3.0;
  else e = mpqc_oracle_reduce_symm(o, v, n_occ, n_frozen, eps, res);
  free(x); free(t3); free(res);
  return e;
}
