/*
 * Minimal C host of the (T) library: shows that include/mpqc_t.h is a plain C ABI (no C++/torch types).
 * Builds with:  gcc -std=c99 -Iinclude examples/c_host.c -o c_host -Lmpqc_b200 -lmpqc_t_cuda -Wl,-rpath,$PWD/mpqc_b200
 * Reads an MPQCT001 dump (mpqc_b200/dump.py, integration/ccsd_t_gpu_impl.h) and prints E(T).
 * Exit code: 0 ok, 2 no CUDA device (mirrors mpqc's FeatureDisabled exit code, mpqc.cpp:261-264), 1 otherwise.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpqc_t.h"

static double* read_doubles(FILE* f, size_t n) {
  double* p = (double*)malloc(n * sizeof(double));
  if (!p || fread(p, sizeof(double), n, f) != n) {
    free(p);
    return NULL;
  }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s problem.mpqct [ngpu]\n%s\n", argv[0], mpqc_t_version());
    return 1;
  }
  FILE* f = fopen(argv[1], "rb");
  char magic[8];
  int64_t hdr[4];
  if (!f || fread(magic, 1, 8, f) != 8 || memcmp(magic, "MPQCT001", 8) != 0 || fread(hdr, sizeof(int64_t), 4, f) != 4) {
    fprintf(stderr, "cannot read dump %s\n", argv[1]);
    return 1;
  }
  const int64_t o = hdr[0], v = hdr[1], nf = hdr[2], nall = hdr[3];
  double* eps = read_doubles(f, (size_t)nall);
  double* t1 = read_doubles(f, (size_t)(v * o));
  double* t2 = read_doubles(f, (size_t)(v * v * o * o));
  double* g_abij = read_doubles(f, (size_t)(v * v * o * o));
  double* g_aijk = read_doubles(f, (size_t)(v * o * o * o));
  double* g_abci = read_doubles(f, (size_t)(v * v * v * o));
  fclose(f);
  if (!eps || !t1 || !t2 || !g_abij || !g_aijk || !g_abci) {
    fprintf(stderr, "truncated dump\n");
    return 1;
  }
  mpqc_t_problem p;
  p.o = o;
  p.v = v;
  p.eps_occ = eps + nf;       /* eps[i + n_frozen]  ccsd_t.h:2306-2311 */
  p.eps_vir = eps + nf + o;   /* eps[a + n_occ] */
  p.t1 = t1;
  p.t2 = t2;
  p.g_abij = g_abij;
  p.g_aijk = g_aijk;
  p.g_abci = g_abci;
  mpqc_t_options opt;
  memset(&opt, 0, sizeof(opt));
  opt.ngpu = argc > 2 ? atoi(argv[2]) : 1;
  opt.unit_count = -1;
  opt.verbose = 1;
  double e_t = 0.0;
  mpqc_t_stats st;
  int rc = mpqc_t_energy(&p, &opt, &e_t, &st);
  if (rc != MPQC_T_OK) {
    fprintf(stderr, "mpqc_t_energy: %s; %s\n", mpqc_t_strerror(rc), mpqc_t_last_error());
    return rc == MPQC_T_ERR_NO_DEVICE ? 2 : 1;
  }
  printf("E(T) = %.15f  units = %lld  launches = %lld\n", e_t, (long long)st.units, (long long)st.kernel_launches);
  free(eps); free(t1); free(t2); free(g_abij); free(g_aijk); free(g_abci);
  return 0;
}
