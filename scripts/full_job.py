"""Whole-job (T): every occupied triple of a named workload through ONE mpqc_t_energy call on host buffers,
one process driving --ngpu devices (static + work-stealing split, NCCL sum).  Prints wall time and TFLOP/s with the
published work model 2 o^3 v^3 (v+o)."""
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
o, v, desc = WORKLOADS[a.workload]
lib = L.load()
pd = make_problem_torch(o, v, "cuda:0")
host = to_host(pd, pin=True)
del pd
torch.cuda.empty_cache()
prob = L.make_problem(o, v, host["eps_occ"], host["eps_vir"], host["t1"], host["t2"], host["g_abij"], host["g_aijk"], host["g_abci"])
opt = L.Options()
opt.ngpu, opt.unit_count, opt.use_nccl, opt.verbose = a.ngpu, -1, a.nccl, 2
e, st = C.c_double(), L.Stats()
t0 = time.perf_counter()
L.check(lib.mpqc_t_energy(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "mpqc_t_energy")
wall = time.perf_counter() - t0
out = dict(workload=a.workload, o=o, v=v, ngpu=a.ngpu, nccl=a.nccl, e_t=e.value, wall_s=wall, units=st.units,
           seconds_upload=st.seconds_upload, seconds_relayout=st.seconds_relayout, seconds_compute=st.seconds_compute,
           tflops_model=lib.mpqc_t_flops(o, v) / wall * 1e-12, tflops_units=st.flops / wall * 1e-12,
           tflops_compute_only=st.flops / st.seconds_compute * 1e-12)
print(json.dumps(out), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"full_job_{a.workload}_n{a.ngpu}.json"), "w") as f:
    json.dump(out, f)
