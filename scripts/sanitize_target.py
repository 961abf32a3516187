"""Developer tool (not product code): uses the oracle only as the checker.
End-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): dense inputs, density-fitted inputs with all
panels resident, and the panel cache (occupied block 1), at sizes that exercise one and several column tiles, the
trailing-fragment skip, the half-filled last k-block and the plain NT-GEMM mode of the W-contraction kernel.

    compute-sanitizer --tool memcheck python scripts/sanitize_target.py 3,17 4,8 6,65 3,130
"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem
from oracle import ccsd_t_oracle as oc
lib = L.load()
sizes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(3, 17), (4, 8)]
worst = 0.0
for (o, v) in sizes:
    p = make_problem(o, v, seed=3)
    ref = oc.ijk_driven(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    for mode in ("dense", "df", "df-panels"):
        opt = L.Options(); opt.ngpu, opt.unit_count = 1, -1
        e, st = C.c_double(), L.Stats()
        if mode == "dense":
            prob = L.make_problem(o, v, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
            L.check(lib.mpqc_t_energy(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "energy")
        else:
            opt.df_block = 1 if mode == "df-panels" else -1
            prob = L.make_df_problem(o, v, p["naux"], p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["x_ab"], p["x_ij"], p["x_ai"])
            L.check(lib.mpqc_t_energy_df(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "energy_df")
        worst = max(worst, abs(e.value - ref))
        print(o, v, mode, e.value, abs(e.value - ref), flush=True)
assert worst < 1e-10, worst
print("sanitize_target: worst |diff| vs oracle", worst)
