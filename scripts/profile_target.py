"""Short profiling target for ncu: N units of a named workload through the split-phase C ABI."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem_torch

o, v, units = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib = L.load()
pd = make_problem_torch(o, v, "cuda")
prob = L.make_problem(o, v, pd["eps_occ"], pd["eps_vir"], pd["t1"], pd["t2"], pd["g_abij"], pd["g_aijk"], pd["g_abci"])
h = C.c_void_p()
L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
L.check(lib.mpqc_t_upload(h, C.byref(prob), 1, None), "upload")
st = L.Stats()
e = C.c_double()
L.check(lib.mpqc_t_run(h, 11, 5, units, batch, C.byref(e), None, C.byref(st)), "run")
print("E", e.value, "compute s", st.seconds_compute, "TF", st.flops / st.seconds_compute * 1e-12)
lib.mpqc_t_destroy(h)
