/*
 * C host of the multi-GPU form of the (T) library (include/mpqc_t.h, "persistent communicator + one collective call").
 *
 *   c_host_comm problem.mpqct [ngpu]              one process drives ngpu devices (local mode)
 *
 * The same three calls serve one MPI rank per GPU (rank mode): rank 0 calls mpqc_t_comm_unique_id, the host program
 * broadcasts the 128 bytes (MPI_Bcast / madness world.gop.broadcast), every rank calls mpqc_t_comm_create_rank -- see
 * the commented block below and scripts/check_rank_mode.py for a complete multi-process host.
 * Reads an MPQCT001 dump into PAGE-LOCKED host buffers (mpqc_t_host_alloc), creates the communicator once, and calls
 * mpqc_t_energy_comm twice (the second call re-uses the device memory the communicator kept).
 * Exit code: 0 ok, 2 no CUDA device (mirrors mpqc's FeatureDisabled exit code, mpqc.cpp:261-264), 1 otherwise.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mpqc_t.h"

static double* read_pinned(FILE* f, size_t n) {
  void* p = NULL;
  if (mpqc_t_host_alloc(&p, n * sizeof(double)) != MPQC_T_OK) return NULL;
  if (fread(p, sizeof(double), n, f) != n) {
    mpqc_t_host_free(p);
    return NULL;
  }
  return (double*)p;
}

static int report(int rc, const char* what) {
  fprintf(stderr, "%s failed: %s; %s\n", what, mpqc_t_strerror(rc), mpqc_t_last_error());
  return rc == MPQC_T_ERR_NO_DEVICE ? 2 : 1;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s problem.mpqct [ngpu]\n%s\n", argv[0], mpqc_t_version());
    return 1;
  }
  if (mpqc_t_device_count() < 1) {
    fprintf(stderr, "%s\n", mpqc_t_strerror(MPQC_T_ERR_NO_DEVICE));
    return 2;
  }
  const int ngpu = argc > 2 ? atoi(argv[2]) : mpqc_t_device_count();
  FILE* f = fopen(argv[1], "rb");
  char magic[8];
  int64_t hdr[4];
  if (!f || fread(magic, 1, 8, f) != 8 || memcmp(magic, "MPQCT001", 8) != 0 || fread(hdr, sizeof(int64_t), 4, f) != 4) {
    fprintf(stderr, "cannot read dump %s\n", argv[1]);
    return 1;
  }
  const int64_t o = hdr[0], v = hdr[1], nf = hdr[2], nall = hdr[3];
  double* eps = read_pinned(f, (size_t)nall);
  mpqc_t_problem p;
  p.o = o;
  p.v = v;
  p.t1 = read_pinned(f, (size_t)(v * o));
  p.t2 = read_pinned(f, (size_t)(v * v * o * o));
  p.g_abij = read_pinned(f, (size_t)(v * v * o * o));
  p.g_aijk = read_pinned(f, (size_t)(v * o * o * o));
  p.g_abci = read_pinned(f, (size_t)(v * v * v * o));
  fclose(f);
  if (!eps || !p.t1 || !p.t2 || !p.g_abij || !p.g_aijk || !p.g_abci) {
    fprintf(stderr, "truncated dump or no page-locked memory\n");
    return 1;
  }
  p.eps_occ = eps + nf;     /* eps[i + n_frozen]  ccsd_t.h:2306-2311 */
  p.eps_vir = eps + nf + o; /* eps[a + n_occ] */

  /* once per wave function: CUDA contexts (in parallel) + NCCL communicator; seconds, so not inside the (T) call */
  mpqc_t_comm* comm = NULL;
  int rc = mpqc_t_comm_create_local(&comm, ngpu, NULL);
  /* rank mode instead (one MPI rank per GPU):
   *   mpqc_t_unique_id id;  if (rank == 0) mpqc_t_comm_unique_id(&id);
   *   MPI_Bcast(&id, sizeof id, MPI_BYTE, 0, MPI_COMM_WORLD);
   *   rc = mpqc_t_comm_create_rank(&comm, nranks, rank, &id, rank % mpqc_t_device_count());            */
  if (rc != MPQC_T_OK) return report(rc, "mpqc_t_comm_create_local");

  mpqc_t_options opt;
  memset(&opt, 0, sizeof(opt));
  opt.unit_count = -1; /* the whole (i >= j >= k) list, sharded over the communicator */
  for (int call = 0; call < 2; ++call) {
    double e_t = 0.0;
    mpqc_t_stats st;
    rc = mpqc_t_energy_comm(comm, &p, &opt, &e_t, &st); /* TOTAL E(T): inputs replicated over NVLink, ncclAllReduce sum */
    if (rc != MPQC_T_OK) {
      mpqc_t_comm_destroy(comm);
      return report(rc, "mpqc_t_energy_comm");
    }
    printf("call %d: E(T) = %.15f  on %d of %d GPUs  (%lld units, upload %.3f s, relayout %.3f s, triples %.3f s, total %.3f s)\n",
           call, e_t, (int)st.ngpu, mpqc_t_comm_size(comm), (long long)st.units, st.seconds_upload, st.seconds_relayout,
           st.seconds_compute, st.seconds_total);
  }
  mpqc_t_comm_destroy(comm);
  mpqc_t_host_free(eps);
  mpqc_t_host_free((void*)p.t1);
  mpqc_t_host_free((void*)p.t2);
  mpqc_t_host_free((void*)p.g_abij);
  mpqc_t_host_free((void*)p.g_aijk);
  mpqc_t_host_free((void*)p.g_abci);
  return 0;
}
