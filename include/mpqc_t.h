/*
 * mpqc_t.h -- C ABI of libmpqc_t_cuda.so: the B200 (sm_100a) implementation of the
 * closed-shell CCSD perturbative-triples (T) energy correction.
 *
 * This is the drop-in boundary for MPQC4's  CCSD_T<Tile,Policy>::compute_ccsd_t()
 *   /root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:144-177  (dispatcher)
 *   /root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:200-711  (coarse, the default)
 *   /root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:1127-1170 (straight)
 * i.e. everything between "T1/T2, orbital energies and the three integral classes are in hand"
 * and "a double comes back".  The MPQC-side adapter that gathers the TiledArray objects into the
 * dense buffers below is integration/ccsd_t_gpu_impl.h, added to the reference's own class by
 * integration/mpqc_ccsd_t_gpu.patch.
 *
 * Plain pointers and sizes only; no C++ or torch types; no exception crosses this boundary.
 * All arrays are IEEE double, dense, row-major (last index fastest), in exactly the layouts
 * the reference's getters produce:
 *
 *   eps_occ[o]          = eps[n_frozen .. n_occ)                 ccsd_t.h:2299-2311, ccsd.h:141-148
 *   eps_vir[v]          = eps[n_occ .. n_all)
 *   t1    [v][o]        t1("a,i")                                ccsd.h:165-179
 *   t2    [v][v][o][o]  t2("a,b,i,j")
 *   g_abij[v][v][o][o]  <ij|ab>  result("a,b,i,j")               ccsd_t.h:2238-2244
 *   g_aijk[v][o][o][o]  <ij|ka>  result("a,i,j,k")               ccsd_t.h:2210-2221
 *   g_abci[v][v][v][o]  <ia|bc>  result("a,b,c,i")               ccsd_t.h:2224-2235
 *
 * Work units.  E(T) = sum over ordered occupied triples i>=j>=k (i==j==k excluded, weight 0) of
 * w_ijk * sum_abc (W+V) Z / D, w = 2 (all distinct) or 1 (two equal).  Triples are enumerated
 * i-major:  for i in [0,o) for j in [0,i] for k in [0,j], skipping i==j==k;  mpqc_t_triple_count(o)
 * = o(o+1)(o+2)/6 - o.  A run processes the units  first, first+stride, ...  (count of them), which is
 * how the path is sharded over GPUs / ranks (replaces the round-robin of ccsd_t.h:477-480).  The
 * partial energies are summed by the caller (replaces gop.sum, ccsd_t.h:692).
 */
#ifndef MPQC_T_H
#define MPQC_T_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPQC_T_ABI_VERSION 2

/* status codes (mapped to mpqc::Exception subclasses by the adapter, util/core/exception.h:93-593) */
enum {
  MPQC_T_OK = 0,
  MPQC_T_ERR_BAD_ARG = 1,    /* -> InputError / ProgrammingError */
  MPQC_T_ERR_NO_DEVICE = 2,  /* -> FeatureDisabled (no CUDA device; there is NO CPU fallback) */
  MPQC_T_ERR_OOM = 3,        /* -> MemAllocFailed */
  MPQC_T_ERR_CUDA = 4,       /* -> ProgrammingError (CUDA runtime / driver error) */
  MPQC_T_ERR_NCCL = 5,       /* -> ProgrammingError (NCCL error) */
  MPQC_T_ERR_INTERNAL = 6
};

typedef struct mpqc_t_problem {
  int64_t o;              /* active occupied   (trange1_engine()->get_active_occ(), trange1_engine.h:60-72) */
  int64_t v;              /* virtuals          (trange1_engine()->get_vir()) */
  const double* eps_occ;  /* [o] */
  const double* eps_vir;  /* [v] */
  const double* t1;       /* [v][o] */
  const double* t2;       /* [v][v][o][o] */
  const double* g_abij;   /* [v][v][o][o] */
  const double* g_aijk;   /* [v][o][o][o] */
  const double* g_abci;   /* [v][v][v][o] */
} mpqc_t_problem;

/* Density-fitted form of the same inputs: instead of the three four-index integral
 * classes the caller hands the three-centre factors the reference's CCSD already holds when is_df() is true,
 *   Xab[K][a][b] = (K|G|a b)[inv_sqr]   CCSD::get_Xab()  ccsd.h:480-483
 *   Xij[K][i][j] = (K|G|i j)[inv_sqr]   CCSD::get_Xij()  ccsd.h:485-488
 *   Xai[K][a][i] = (K|G|a i)[inv_sqr]   CCSD::get_Xai()  ccsd.h:490-493
 * and the library forms  <ij|ab> = sum_K Xai[K,a,i] Xai[K,b,j],  <ij|ka> = sum_K Xij[K,i,k] Xai[K,a,j],
 * <ia|bc> = sum_K Xai[K,b,i] Xab[K,a,c]  directly in their occupied-major device layouts (what the [df] getters
 * of ccsd_t.h:2210-2244 evaluate on the host).  The v^3 o tensor never exists on the host or crosses PCIe. */
typedef struct mpqc_t_df_problem {
  int64_t o, v, naux;
  const double* eps_occ;  /* [o] */
  const double* eps_vir;  /* [v] */
  const double* t1;       /* [v][o] */
  const double* t2;       /* [v][v][o][o] */
  const double* x_ab;     /* [naux][v][v] */
  const double* x_ij;     /* [naux][o][o] */
  const double* x_ai;     /* [naux][v][o] */
} mpqc_t_df_problem;

typedef struct mpqc_t_options {
  int32_t ngpu;               /* number of devices this process drives (>=1); 0 -> 1 */
  const int32_t* device_ids;  /* [ngpu] CUDA ordinals, NULL -> 0..ngpu-1 */
  int32_t verbose;            /* 0 silent, 1 prints the reference's "(T) Energy: ... Time: ... S" line */
  int32_t inputs_on_device;   /* 0: problem pointers are host memory; 1: device memory on device_ids[0] (ngpu must be 1) */
  int64_t unit_first;         /* first triple unit of this process' shard (rank) */
  int64_t unit_stride;        /* stride between this process' units (world size); 0 -> 1 */
  int64_t unit_count;         /* number of units to process; <0 -> all remaining with that stride */
  int32_t batch;              /* triples per kernel launch; 0 -> auto */
  int32_t steal_chunk;        /* in-process multi-GPU: triples per work-stealing grab; 0 -> auto */
  int32_t use_nccl;           /* mpqc_t_energy with ngpu > 1 and no communicator: 1 = build a communicator for this call
                                 (NVLink input replication + ncclAllReduce sum), 0 = replicated uploads + host sum */
  int32_t df_block;           /* density-fitted inputs: 0 = automatic (all operand panels resident when they fit, else a panel
                                 cache), -1 = resident, b > 0 = panel cache walking occupied blocks of edge b (3 b slots) */
  int32_t reserved[4];
} mpqc_t_options;

typedef struct mpqc_t_stats {
  double seconds_total;     /* wall time of the call */
  double seconds_upload;    /* host -> device copies */
  double seconds_relayout;  /* integral/amplitude blocking on device (replaces ccsd_t.h:2219,2233,2242 + reblock) */
  double seconds_compute;   /* triples loop, device-timed with CUDA events (max over devices) */
  double seconds_contract;  /* share of seconds_compute spent in the W contraction kernel (single-GPU profile mode only, else 0) */
  double seconds_energy;    /* share spent in the fused V/symmetrise/denominator/reduce kernel (same) */
  double flops;             /* algorithmic: 12 v^3 (v+o) per triple processed */
  double flops_executed;    /* including tile padding */
  int64_t units;            /* triples processed by this call */
  int64_t kernel_launches;  /* number of kernels launched by this call */
  int64_t bytes_h2d;
  int64_t bytes_d2h;
  int32_t ngpu;
  int32_t reserved[7];
} mpqc_t_stats;

typedef struct mpqc_t_handle mpqc_t_handle; /* opaque: one device's resident, re-laid-out problem */
typedef struct mpqc_t_comm mpqc_t_comm;     /* opaque: the GPUs that share one (T) job, and their NCCL communicator */
typedef struct mpqc_t_unique_id { char internal[128]; } mpqc_t_unique_id;   /* == ncclUniqueId */

/* ---- one-shot entry point: what the adapter's compute_ccsd_t() calls ------------------------ */
/* Computes this process' partial E(T) over its units (all units when unit_first=0, unit_stride=1,
 * unit_count<0).  Caller owns every buffer for the duration of the call; nothing is retained. */
int mpqc_t_energy(const mpqc_t_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats);

/* same contract, density-fitted inputs */
int mpqc_t_energy_df(const mpqc_t_df_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats);

/* ---- multi-GPU form: a persistent communicator + one collective call ------------------------------------
 * Replaces, inside the library, the reference's replicated integrals and its final  gop.sum  (ccsd_t.h:692):
 *   - every rank calls mpqc_t_energy_comm with the SAME host problem and options; the job (unit_first, unit_stride,
 *     unit_count of the options: all units by default) is sharded over the ranks of the communicator
 *     (opt.ngpu / device_ids are ignored: the communicator names the devices);
 *   - each host tensor crosses PCIe once in total: rank r copies 1/N of it over its own link and one ncclAllGather
 *     over NVLink completes it on every GPU;
 *   - the per-unit energies are summed by one ncclAllReduce and *e_t is the TOTAL E(T) of the job on every rank,
 *     bit-identical to the single-GPU result.
 * Two launch modes:
 *   rank mode  - one MPI rank per GPU: rank 0 calls mpqc_t_comm_unique_id, the host program broadcasts the 128
 *                bytes (world.gop.broadcast), every rank calls mpqc_t_comm_create_rank;
 *   local mode - one process drives ngpu devices (one host thread per GPU inside the call, joined before it
 *                returns; static share + work-stealing tail): mpqc_t_comm_create_local.
 * Creating a communicator creates the CUDA contexts (in parallel) and the NCCL communicator, which takes seconds:
 * do it once, e.g. when the wave function object is constructed, not per (T) call.  Errors are agreed upon
 * collectively: if any rank fails (e.g. out of device memory) every rank returns an error instead of blocking. */
int mpqc_t_comm_unique_id(mpqc_t_unique_id* id);
int mpqc_t_comm_create_rank(mpqc_t_comm** c, int32_t nranks, int32_t rank, const mpqc_t_unique_id* id, int32_t device);
int mpqc_t_comm_create_local(mpqc_t_comm** c, int32_t ngpu, const int32_t* device_ids /* NULL -> 0..ngpu-1 */);
int mpqc_t_comm_size(const mpqc_t_comm* c);
/* The communicator keeps the device memory of the last problem (operand panels, workspaces) for the next call with the
 * same (o, v) -- allocating and freeing tens of GB costs seconds on some hosts; this hands it back to the driver now. */
int mpqc_t_comm_release_cache(mpqc_t_comm* c);
int mpqc_t_comm_destroy(mpqc_t_comm* c);
int mpqc_t_energy_comm(mpqc_t_comm* c, const mpqc_t_problem* p, const mpqc_t_options* opt, double* e_t, mpqc_t_stats* stats);
int mpqc_t_energy_df_comm(mpqc_t_comm* c, const mpqc_t_df_problem* p, const mpqc_t_options* opt, double* e_t,
                          mpqc_t_stats* stats);

/* page-locked host memory for the dense buffers the adapter gathers into (full PCIe speed for the uploads) */
int mpqc_t_host_alloc(void** ptr, size_t bytes);
int mpqc_t_host_free(void* ptr);

/* ---- split-phase API (lets uploads be timed separately; used by bench.py and the tests) ------
 * Device-resident inputs (on_device = 1) may have been produced on any stream of the caller: upload starts with a
 * cudaDeviceSynchronize(), so no event hand-off is needed; the handle then works on its own non-blocking stream. */
int mpqc_t_create(mpqc_t_handle** h, int64_t o, int64_t v, int32_t device);
/* host (on_device=0) or device (on_device=1) buffers -> occupied-major operand layouts in HBM */
int mpqc_t_upload(mpqc_t_handle* h, const mpqc_t_problem* p, int32_t on_device, mpqc_t_stats* stats);
/* density-fitted inputs -> the same device layouts; the integral classes are assembled on the device by the library's
 * own TMA + DMMA GEMM pipeline (no library GEMM), all panels at once or on demand (mpqc_t_set_df_block) */
int mpqc_t_upload_df(mpqc_t_handle* h, const mpqc_t_df_problem* p, int32_t on_device, mpqc_t_stats* stats);
/* process units first, first+stride, ... (count of them; <0 = to the end).  partial_e = weighted sum
 * over those units (summed in unit order, so any sharding gives bit-identical per-unit terms);
 * unit_e (optional, may be NULL) receives the weighted per-unit energies [count].  batch 0 = auto. */
int mpqc_t_run(mpqc_t_handle* h, int64_t first, int64_t stride, int64_t count, int32_t batch,
               double* partial_e, double* unit_e, mpqc_t_stats* stats);
/* collective form of mpqc_t_run for one rank per GPU: (first, stride, count) name the JOB, every rank of the
 * rank-mode communicator runs job positions rank, rank+R, ... on its own resident handle, and one ncclAllReduce on the
 * handle's stream (device buffers; no host round trip before the collective) completes the job-length unit-energy
 * vector on every rank.  total_e = E(T) of the whole job, identical on all ranks and for any R; unit_e (optional)
 * receives the [count] unit energies of the job. */
int mpqc_t_run_comm(mpqc_t_handle* h, mpqc_t_comm* c, int64_t first, int64_t stride, int64_t count, int32_t batch,
                    double* total_e, double* unit_e, mpqc_t_stats* stats);
/* same as mpqc_t_run, and additionally the decomposition of the SAME energy over virtual-block triples:
 * vblock_e[tt], tt = 0 .. energy_tile_sets-1 (mpqc_t_plan), enumerates 8-wide virtual tiles TA >= TB >= TC in the
 * order of the reference's coarse loop with block size 8 (global_iter - 1, ccsd_t.h:443-480).  Run over ALL units,
 * vblock_e[tt] is the energy that loop iteration contributes (ccsd_t.h:619-638) and sum_tt vblock_e[tt] = E(T):
 * the whole-job result can be checked against sampled iterations of the reference algorithm. */
int mpqc_t_run_vblocks(mpqc_t_handle* h, int64_t first, int64_t stride, int64_t count, int32_t batch,
                       double* partial_e, double* unit_e, double* vblock_e, mpqc_t_stats* stats);
/* The W build alone, batched: W^{abc}_{ijk} (the six particle + six hole contractions, ccsd_t.h:1142-1146) for n
 * ARBITRARY occupied triples (no ordering required), written as dense [n][v][v][v] arrays (a,b,c row-major) into device
 * memory of the handle's device (out_on_device = 1) or host memory.  This is the entry point an iterative-triples
 * model (CC3 / CCSDT-1, cc3.h:55+, ccsdt1.h:55+: the same contraction shapes every CC iteration) would call. */
int mpqc_t_w_batch(mpqc_t_handle* h, const int32_t* triples /* [n][3] */, int64_t n, double* w_out, int32_t out_on_device);
/* debugging / parity aid: W^{abc}_{ijk} of one occupied triple as a dense [v][v][v] host array */
int mpqc_t_debug_w(mpqc_t_handle* h, int32_t i, int32_t j, int32_t k, double* w_host);
/* what the last upload decided for this handle */
enum {
  MPQC_T_QUERY_PANEL_SLOTS = 0,   /* operand panels held on the device (o when resident) */
  MPQC_T_QUERY_PANEL_MODE = 1,    /* 1: panel cache (panels built on demand), 0: all resident */
  MPQC_T_QUERY_FLAT = 2,          /* 1: flat rows + transposed operand copy, 0: row patches */
  MPQC_T_QUERY_PANELS_BUILT = 3,  /* operand panels assembled from three-centre factors so far */
  MPQC_T_QUERY_PANEL_BLOCK = 4    /* occupied block edge of the panel walk */
};
int mpqc_t_query(mpqc_t_handle* h, int32_t what, int64_t* value);
/* density-fitted inputs, split-phase form of mpqc_t_options.df_block: call before mpqc_t_upload_df */
int mpqc_t_set_df_block(mpqc_t_handle* h, int32_t block);
/* CUDA stream the handle launches on (cudaStream_t as void*), for event timing by the caller */
void* mpqc_t_stream(mpqc_t_handle* h);
int mpqc_t_destroy(mpqc_t_handle* h);

/* ---- helpers ------------------------------------------------------------------------------- */
int64_t mpqc_t_triple_count(int64_t o);
/* unit index -> (i,j,k) of the enumeration above; returns MPQC_T_ERR_BAD_ARG when out of range */
int mpqc_t_triple_of_unit(int64_t o, int64_t unit, int32_t* i, int32_t* j, int32_t* k);
/* Host-only: which positions of the job (first, first+stride, ...; count of them, <0 = all) worker `rank` of `nranks` runs
 * statically -- exactly what mpqc_t_energy[_comm] does inside.  panel_block = 0: unit-cyclic (rank, rank+nranks, ...);
 * when all workers share a process (all_local) only the first 7/8 are static and *tail_begin is where the work-stealing
 * tail starts.  panel_block > 0 (operand panel cache): occupied-block groups dealt largest-first to the least loaded
 * worker.  Returns the number of positions (written to positions[0..capacity)), or -1 on bad arguments. */
int64_t mpqc_t_shard_plan(int64_t o, int64_t first, int64_t stride, int64_t count, int32_t nranks, int32_t rank,
                          int32_t all_local, int32_t panel_block, int64_t* positions, int64_t capacity, int64_t* tail_begin);
double mpqc_t_flops(int64_t o, int64_t v);           /* 2 o^3 v^3 (v+o), the published work model */
double mpqc_t_unit_flops(int64_t o, int64_t v);      /* 12 v^3 (v+o) */
int mpqc_t_device_count(void);
const char* mpqc_t_version(void);
const char* mpqc_t_strerror(int status);
const char* mpqc_t_last_error(void);                 /* thread-local detail string of the last failure */

/* Host-only: the tiling plan the W-contraction kernel would use for (o, v) -- row mode, tile counts, column-fragment
 * count -- and the resulting fraction of executed tensor-core FLOPs that are algorithmic.  No device is touched. */
typedef struct mpqc_t_plan_info {
  int64_t kp;            /* padded contraction length roundup8(v+o) (>= 16) */
  int32_t flat;          /* 1: flattened (p,q) rows + transposed operand copy; 0: (tp x tq) row patches */
  int32_t tp, tq;        /* row patch (patch mode) */
  int32_t row_tiles;     /* 128-row tiles per group */
  int32_t col_tiles;     /* column tiles */
  int32_t nfrag;         /* 8-column fragments per column tile (kernel instantiation) */
  int32_t skip_last;     /* 1: the last column tile drops its trailing fragment */
  int32_t energy_tile_sets; /* blocks per triple of the energy kernel */
  double flop_efficiency;   /* 12 v^3 (v+o) / executed FLOPs per triple */
  double bytes_operands;    /* resident operand bytes (A [+AT], B, GV) */
} mpqc_t_plan_info;
int mpqc_t_plan(int64_t o, int64_t v, int32_t flat, mpqc_t_plan_info* out);

/* Host-only: device-memory model of the density-fitted path.  block = 0: every operand panel
 * resident; block = b > 0: panel cache of 3 b slots -- the v^3 o operand is never resident, panels A_x are built on demand
 * from the three-centre factors on the library's own TMA + DMMA pipeline while the units are walked occupied-block-wise. */
typedef struct mpqc_t_df_plan_info {
  int32_t npanel, block, panel_mode, flat;
  double bytes_panels, bytes_b, bytes_gv, bytes_t2, bytes_factors, bytes_w_workspace, bytes_total;
  double build_flop_fraction;   /* FLOPs spent building panels / FLOPs of the triples, whole job */
} mpqc_t_df_plan_info;
int mpqc_t_plan_df(int64_t o, int64_t v, int64_t naux, int32_t block, int32_t flat, mpqc_t_df_plan_info* out);

/* FP64 pipe microbenchmarks used to fix the roofline denominator on the box:
 * which = 0: DMMA.8x8x4 issue-bound loop (32 warps/SM), 1: DFMA issue-bound loop, 2: DMMA loop at the W-contraction
 * kernel's occupancy (8 warps/SM, two per scheduler).  Returns TFLOP/s in *tflops. */
int mpqc_t_microbench(int32_t device, int32_t which, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MPQC_T_H */
