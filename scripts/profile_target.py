"""Short profiling target for ncu: N units of a named workload through the split-phase C ABI.

    python scripts/profile_target.py O V UNITS [BATCH] [dense|df|dfpanel]

dense: device-resident dense inputs (the W-contraction + energy kernels); df: density-fitted upload with every panel
resident (adds the plain NT-GEMM launches of the same kernel that assemble the integrals); dfpanel: panel cache, block 4."""
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
mode = sys.argv[5] if len(sys.argv) > 5 else "dense"
lib = L.load()
pd = make_problem_torch(o, v, "cuda", dense_abci=(mode == "dense"))
h = C.c_void_p()
L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
up = L.Stats()
if mode == "dense":
    prob = L.make_problem(o, v, pd["eps_occ"], pd["eps_vir"], pd["t1"], pd["t2"], pd["g_abij"], pd["g_aijk"], pd["g_abci"])
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 1, C.byref(up)), "upload")
else:
    prob = L.make_df_problem(o, v, int(pd["naux"]), pd["eps_occ"], pd["eps_vir"], pd["t1"], pd["t2"], pd["x_ab"], pd["x_ij"], pd["x_ai"])
    L.check(lib.mpqc_t_set_df_block(h, 4 if mode == "dfpanel" else -1), "set_df_block")
    L.check(lib.mpqc_t_upload_df(h, C.byref(prob), 1, C.byref(up)), "upload_df")
st = L.Stats()
e = C.c_double()
L.check(lib.mpqc_t_run(h, 11, 5, units, batch, C.byref(e), None, C.byref(st)), "run")
print("mode", mode, "E", e.value, "upload+build s", up.seconds_relayout, "compute s", st.seconds_compute, "TF",
      st.flops / st.seconds_compute * 1e-12, "launches", up.kernel_launches + st.kernel_launches)
lib.mpqc_t_destroy(h)
