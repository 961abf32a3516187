"""Whole-job (T): every occupied triple of a named workload through ONE mpqc_t_energy_comm call on pinned host buffers.

  python scripts/full_job.py --ngpu N                         one process driving N devices (local communicator:
                                                              threads, NVLink input replication, ncclAllReduce sum)
  torchrun --nproc-per-node N scripts/full_job.py            one rank per GPU (rank-mode communicator of the library;
                                                              the NCCL id is broadcast through gloo)
  ... --df                                                    density-fitted hand-off (mpqc_t_energy_df_comm)

The communicator is created BEFORE the timed region (an MPQC run creates it when the wave function is constructed, long
before (T) starts); its set-up time is reported separately.  Prints wall time of the call (barrier to total energy, max
over ranks) and TFLOP/s with the work model 2 o^3 v^3 (v+o)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem_torch
from bench import WORKLOADS, to_host

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="uracil-trimer-6-31Gs")
ap.add_argument("--ngpu", type=int, default=1)
ap.add_argument("--df", action="store_true")
ap.add_argument("--df-block", type=int, default=0)
ap.add_argument("--verbose", type=int, default=0)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
o, v, desc = WORKLOADS[a.workload]
lib = L.load()
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
pd = make_problem_torch(o, v, f"cuda:{local}", dense_abci=not a.df)
host = to_host(pd, pin=True)
del pd
torch.cuda.empty_cache()
torch.cuda.synchronize()
if a.df:
    prob = L.make_df_problem(o, v, int(host["naux"]), host["eps_occ"], host["eps_vir"], host["t1"], host["t2"],
                             host["x_ab"], host["x_ij"], host["x_ai"])
    call = lib.mpqc_t_energy_df_comm
else:
    prob = L.make_problem(o, v, host["eps_occ"], host["eps_vir"], host["t1"], host["t2"], host["g_abij"], host["g_aijk"],
                          host["g_abci"])
    call = lib.mpqc_t_energy_comm
comm = C.c_void_p()
t0 = time.perf_counter()
if world > 1:
    uid = L.UniqueId()
    if rank == 0:
        L.check(lib.mpqc_t_comm_unique_id(C.byref(uid)), "unique_id")
    t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), t.numpy().tobytes(), 128)
    L.check(lib.mpqc_t_comm_create_rank(C.byref(comm), world, rank, C.byref(uid), local), "comm_create_rank")
else:
    L.check(lib.mpqc_t_comm_create_local(C.byref(comm), a.ngpu, None), "comm_create_local")
setup_s = time.perf_counter() - t0
opt = L.Options()
opt.unit_count, opt.df_block, opt.verbose = -1, a.df_block, (a.verbose if rank == 0 else 0)
if world > 1:
    dist.barrier()
e, st = C.c_double(), L.Stats()
t0 = time.perf_counter()
L.check(call(comm, C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "mpqc_t_energy_comm")
wall = time.perf_counter() - t0
lib.mpqc_t_comm_destroy(comm)
vals = torch.tensor([wall, st.seconds_compute, st.seconds_upload, st.seconds_relayout], dtype=torch.float64)
if world > 1:
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
wall, comp, up, rel = (float(x) for x in vals)
ngpu = world if world > 1 else a.ngpu
if rank == 0:
    out = dict(workload=a.workload, o=o, v=v, ngpu=ngpu, mode="one rank per GPU" if world > 1 else "one process",
               inputs="density-fitted" if a.df else "dense", df_block=a.df_block, e_t=e.value, wall_s=wall,
               seconds_comm_setup=setup_s, seconds_compute_max=comp, seconds_upload_max=up, seconds_relayout_max=rel,
               units=lib.mpqc_t_triple_count(o), tflops_model=lib.mpqc_t_flops(o, v) / wall * 1e-12)
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = ("ranks" if world > 1 else "n") + str(ngpu) + ("_df" if a.df else "") + (f"_b{a.df_block}" if a.df_block else "")
    with open(os.path.join(ROOT, "gpurun_out", f"r02_full_job_{a.workload}_{tag}.json"), "w") as f:
        json.dump(out, f)
if world > 1:
    dist.destroy_process_group()
