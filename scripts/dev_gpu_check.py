"""Developer GPU check (run under gpurun): microbenchmarks, small parity vs the oracle, one timing."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem
from oracle import ccsd_t_oracle as oc

lib = L.load()
out = {}


def micro():
    for which, name in [(0, "dmma"), (1, "dfma")]:
        tf = C.c_double()
        L.check(lib.mpqc_t_microbench(0, which, C.byref(tf)), "microbench")
        out[name + "_tflops"] = tf.value
        print(name, tf.value, flush=True)
    import torch
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / (best * 1e-3) * 1e-12
    print("cublas dgemm", out["cublas_dgemm_8192_tflops"], flush=True)


def parity(o, v, check_w=True):
    p = make_problem(o, v, scale=1.0)
    args = (p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    prob = L.make_problem(o, v, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
    h = C.c_void_p()
    L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
    st = L.Stats()
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 0, C.byref(st)), "upload")
    if check_w:
        for (i, j, k) in [(o - 1, 1, 0), (2, 2, 1), (o - 1, o - 2, o - 2)]:
            w = np.zeros((v, v, v))
            L.check(lib.mpqc_t_debug_w(h, i, j, k, w.ctypes.data_as(L.c_double_p)), "debug_w")
            wr = oc.w_ijk(p["t2"], p["g_aijk"], p["g_abci"], i, j, k)
            print(f"  W({i},{j},{k}) max|diff| = {np.abs(w - wr).max():.3e}  max|W| = {np.abs(wr).max():.3e}", flush=True)
    nt = lib.mpqc_t_triple_count(o)
    e = C.c_double()
    ue = np.zeros(nt)
    L.check(lib.mpqc_t_run(h, 0, 1, -1, 0, C.byref(e), ue.ctypes.data_as(L.c_double_p), C.byref(st)), "run")
    eref, parts = oc.ijk_driven(*args, return_parts=True)
    print(f"o={o} v={v}: E_gpu={e.value:.15e} E_oracle={eref:.15e} diff={abs(e.value - eref):.3e} "
          f"max unit diff={np.abs(ue - parts).max():.3e}", flush=True)
    lib.mpqc_t_destroy(h)
    return abs(e.value - eref)


def timing(o, v, nunits):
    import torch
    from mpqc_b200.synthetic import make_problem_torch
    t0 = time.time()
    p = make_problem_torch(o, v, "cuda")
    torch.cuda.synchronize()
    print("generated synthetic on device in", time.time() - t0, flush=True)
    prob = L.make_problem(o, v, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
    h = C.c_void_p()
    L.check(lib.mpqc_t_create(C.byref(h), o, v, 0), "create")
    st = L.Stats()
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 1, C.byref(st)), "upload")
    print("relayout s:", st.seconds_relayout, flush=True)
    os.environ["MPQC_T_PROFILE"] = "1"
    for rep in range(2):
        st = L.Stats()
        e = C.c_double()
        L.check(lib.mpqc_t_run(h, 0, 7, nunits, 0, C.byref(e), None, C.byref(st)), "run")
        tf = st.flops / st.seconds_compute * 1e-12
        print(f"o={o} v={v} units={st.units} compute={st.seconds_compute:.4f}s contract={st.seconds_contract:.4f}s "
              f"energy={st.seconds_energy:.4f}s  TFLOP/s={tf:.2f} (executed {st.flops_executed / st.seconds_compute * 1e-12:.2f}) "
              f"contract-only TFLOP/s={st.flops / max(st.seconds_contract, 1e-9) * 1e-12:.2f} E={e.value:.12e}", flush=True)
        out[f"timing_{o}_{v}"] = dict(tflops=tf, compute=st.seconds_compute, contract=st.seconds_contract,
                                      energy=st.seconds_energy, units=st.units)
    lib.mpqc_t_destroy(h)


if __name__ == "__main__":
    what = sys.argv[1:] or ["micro", "parity", "timing"]
    if "micro" in what:
        micro()
    if "parity" in what:
        for (o, v) in [(4, 8), (5, 19), (3, 33), (6, 40)]:
            parity(o, v)
    if "timing" in what:
        timing(21, 93, 256)
        timing(63, 297, 24)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dev_check.json"), "w") as f:
        json.dump(out, f, indent=1)
