"""Timing sweep over problem shapes (few units each)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem_torch
lib = L.load()
os.environ["MPQC_T_PROFILE"] = "1"
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for (o, v, units) in shapes:
    pd = make_problem_torch(o, v, "cuda")
    prob = L.make_problem(o, v, pd["eps_occ"], pd["eps_vir"], pd["t1"], pd["t2"], pd["g_abij"], pd["g_aijk"], pd["g_abci"])
    h = C.c_void_p()
    L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
    up = L.Stats()
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 1, C.byref(up)), "upload")
    for rep in range(2):
        st = L.Stats(); e = C.c_double()
        L.check(lib.mpqc_t_run(h, 0, 3, units, 0, C.byref(e), None, C.byref(st)), "run")
    print(f"o={o} v={v} units={st.units}: {st.flops/st.seconds_compute*1e-12:.2f} TF  (executed {st.flops_executed/st.seconds_compute*1e-12:.2f})  "
          f"contract {st.flops/st.seconds_contract*1e-12:.2f} TF (executed {st.flops_executed/st.seconds_contract*1e-12:.2f})  "
          f"energy/unit {st.seconds_energy/st.units*1e6:.1f} us  energy share {st.seconds_energy/st.seconds_compute*100:.1f}%  relayout {up.seconds_relayout:.3f}s", flush=True)
    lib.mpqc_t_destroy(h)
    del pd
    torch.cuda.empty_cache()
