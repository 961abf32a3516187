"""Timing sweep over problem shapes and batch sizes (few units each):  python scripts/sweep.py O,V,UNITS[,BATCH[,BATCH...]] ..."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem_torch
lib = L.load()
os.environ["MPQC_T_PROFILE"] = "1"
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for (o, v, units, *batches) in shapes:
    pd = make_problem_torch(o, v, "cuda")
    prob = L.make_problem(o, v, pd["eps_occ"], pd["eps_vir"], pd["t1"], pd["t2"], pd["g_abij"], pd["g_aijk"], pd["g_abci"])
    h = C.c_void_p()
    L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
    up = L.Stats()
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 1, C.byref(up)), "upload")
    first = max(0, (lib.mpqc_t_triple_count(o) - units) // 2)
    for batch in (batches or [0]):
        for rep in range(3):
            st = L.Stats(); e = C.c_double()
            L.check(lib.mpqc_t_run(h, first, 1, units, batch, C.byref(e), None, C.byref(st)), "run")
        print(f"o={o} v={v} units={st.units} batch={batch}: {st.flops/st.seconds_compute*1e-12:.2f} TF  "
              f"contract {st.flops/st.seconds_contract*1e-12:.2f} TF  energy/unit {st.seconds_energy/st.units*1e6:.1f} us  "
              f"energy share {st.seconds_energy/st.seconds_compute*100:.1f}%  launches {st.kernel_launches}", flush=True)
    lib.mpqc_t_destroy(h)
    del pd
    torch.cuda.empty_cache()
