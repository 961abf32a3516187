"""Whole-job (T): every occupied triple of a named workload through ONE mpqc_t_energy call per process on host buffers.

  python scripts/full_job.py --ngpu N                         one process driving N devices (threads, NCCL sum)
  torchrun --nproc-per-node N scripts/full_job.py            one rank per GPU (units rank, rank+N, ...; NCCL all_reduce)

Prints wall time (barrier to summed energy, max over ranks) and TFLOP/s with the work model 2 o^3 v^3 (v+o)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem_torch
from bench import WORKLOADS, to_host

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="uracil-trimer-6-31Gs")
ap.add_argument("--ngpu", type=int, default=1)
ap.add_argument("--nccl", type=int, default=1)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
o, v, desc = WORKLOADS[a.workload]
lib = L.load()
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pd = make_problem_torch(o, v, f"cuda:{local}")
host = to_host(pd, pin=True)
del pd
torch.cuda.empty_cache()
prob = L.make_problem(o, v, host["eps_occ"], host["eps_vir"], host["t1"], host["t2"], host["g_abij"], host["g_aijk"], host["g_abci"])
opt = L.Options()
opt.unit_count, opt.use_nccl, opt.verbose = -1, a.nccl, (2 if rank == 0 else 0)
if world > 1:
    ids = (C.c_int32 * 1)(local)
    opt.ngpu, opt.device_ids, opt.unit_first, opt.unit_stride = 1, ids, rank, world
    t = torch.zeros(1, dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t)                      # warm the communicator (an MPQC run has its communicator up already)
    dist.barrier()
else:
    opt.ngpu = a.ngpu
torch.cuda.synchronize()
e, st = C.c_double(), L.Stats()
t0 = time.perf_counter()
L.check(lib.mpqc_t_energy(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "mpqc_t_energy")
e_t = e.value
if world > 1:
    t = torch.tensor([e.value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t)                      # the path's one collective (gop.sum, ccsd_t.h:692)
    e_t = float(t.cpu()[0])
wall = time.perf_counter() - t0
if world > 1:
    tw = torch.tensor([wall, st.seconds_compute, st.seconds_upload + st.seconds_relayout], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    wall, comp, up = (float(x) for x in tw.cpu())
else:
    comp, up = st.seconds_compute, st.seconds_upload + st.seconds_relayout
ngpu = world if world > 1 else a.ngpu
if rank == 0:
    out = dict(workload=a.workload, o=o, v=v, ngpu=ngpu, mode="one rank per GPU" if world > 1 else "one process",
               e_t=e_t, wall_s=wall, seconds_compute_max=comp, seconds_upload_relayout_max=up,
               tflops_model=lib.mpqc_t_flops(o, v) / wall * 1e-12)
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = "ranks" if world > 1 else "n"
    with open(os.path.join(ROOT, "gpurun_out", f"full_job_{a.workload}_{tag}{ngpu}.json"), "w") as f:
        json.dump(out, f)
if world > 1:
    dist.destroy_process_group()
