"""Developer tool (not product code): uses the oracle only as the checker."""
"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem
from oracle import ccsd_t_oracle as oc
lib = L.load()
for (o, v) in [(3, 17), (4, 8)]:
    p = make_problem(o, v, seed=3)
    for df in (0, 1):
        opt = L.Options(); opt.ngpu, opt.unit_count = 1, -1
        e, st = C.c_double(), L.Stats()
        if df:
            prob = L.make_df_problem(o, v, p["naux"], p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["x_ab"], p["x_ij"], p["x_ai"])
            L.check(lib.mpqc_t_energy_df(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "energy_df")
        else:
            prob = L.make_problem(o, v, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
            L.check(lib.mpqc_t_energy(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "energy")
        ref = oc.ijk_driven(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
        print(o, v, "df" if df else "dense", e.value, abs(e.value - ref))
