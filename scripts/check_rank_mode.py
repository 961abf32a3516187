"""One rank per GPU through the LIBRARY's communicator (no torch NCCL involved): mpqc_t_comm_unique_id on rank 0, the
128 bytes broadcast by the host program (gloo here, world.gop.broadcast in MPQC), mpqc_t_comm_create_rank on every
rank, then mpqc_t_energy_comm / mpqc_t_energy_df_comm on identical host inputs.  Checks on every rank: total E(T)
identical everywhere, bit-identical to the single-GPU call, within 1e-10 Eh of the CPU oracle; input bytes cross
PCIe once in total.

    torchrun --nnodes 1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/check_rank_mode.py
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from mpqc_b200 import lib as L
from mpqc_b200.synthetic import make_problem
from oracle import ccsd_t_oracle as oc

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("gloo")
lib = L.load()
uid = L.UniqueId()
if rank == 0:
    L.check(lib.mpqc_t_comm_unique_id(C.byref(uid)), "unique_id")
t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
dist.broadcast(t, 0)
C.memmove(C.byref(uid), t.numpy().tobytes(), 128)
comm = C.c_void_p()
L.check(lib.mpqc_t_comm_create_rank(C.byref(comm), world, rank, C.byref(uid), local), "comm_create_rank")

out = {"world": world}
for (o, v) in [(9, 41), (12, 70)]:
    p = make_problem(o, v, seed=300 + v)
    prob = L.make_problem(o, v, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
    dfp = L.make_df_problem(o, v, p["naux"], p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["x_ab"], p["x_ij"], p["x_ai"])
    opt = L.Options()
    opt.unit_count = -1
    e, st = C.c_double(), L.Stats()
    L.check(lib.mpqc_t_energy_comm(comm, C.byref(prob), C.byref(opt), C.byref(e), C.byref(st)), "energy_comm")
    ed, sd = C.c_double(), L.Stats()
    L.check(lib.mpqc_t_energy_df_comm(comm, C.byref(dfp), C.byref(opt), C.byref(ed), C.byref(sd)), "energy_df_comm")
    # density-fitted hand-off with the operand panel cache (occupied block 2): groups of units, not single units, are
    # dealt to the ranks
    opt_p = L.Options()
    opt_p.unit_count, opt_p.df_block = -1, 2
    ep, sp = C.c_double(), L.Stats()
    L.check(lib.mpqc_t_energy_df_comm(comm, C.byref(dfp), C.byref(opt_p), C.byref(ep), C.byref(sp)), "energy_df_comm(panels)")
    # single-GPU plain call on this rank's device
    o1 = L.Options()
    ids = (C.c_int32 * 1)(local)
    o1.ngpu, o1.device_ids, o1.unit_count = 1, ids, -1
    e1, s1 = C.c_double(), L.Stats()
    L.check(lib.mpqc_t_energy(C.byref(prob), C.byref(o1), C.byref(e1), C.byref(s1)), "energy")
    # split-phase collective run on a resident handle
    h = C.c_void_p()
    L.check(lib.mpqc_t_create(C.byref(h), o, v, local), "create")
    L.check(lib.mpqc_t_upload(h, C.byref(prob), 0, None), "upload")
    er, sr = C.c_double(), L.Stats()
    n = lib.mpqc_t_triple_count(o)
    ue = np.zeros(n)
    L.check(lib.mpqc_t_run_comm(h, comm, 0, 1, -1, 0, C.byref(er), ue.ctypes.data_as(L.c_double_p), C.byref(sr)), "run_comm")
    lib.mpqc_t_destroy(h)
    e_ref = oc.ijk_driven(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    allv = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    units_panel = torch.tensor([float(sp.units)], dtype=torch.float64)
    dist.all_reduce(units_panel)
    assert int(units_panel[0]) == n, (int(units_panel[0]), n)      # every unit ran exactly once across the ranks
    dist.all_gather(allv, torch.tensor([e.value, ed.value, er.value], dtype=torch.float64))
    same = all(bool(torch.equal(x, allv[0])) for x in allv)
    h2d = torch.tensor([float(st.bytes_h2d)], dtype=torch.float64)
    dist.all_reduce(h2d)
    dense_bytes = 8 * sum(p[k].size for k in ("t2", "g_abij", "g_aijk", "g_abci"))
    rec = {"o": o, "v": v, "e_comm": e.value, "e_single": e1.value, "e_df_comm": ed.value, "e_run_comm": er.value,
           "oracle": e_ref, "identical_on_all_ranks": same, "bitwise_equal_single_gpu": e.value == e1.value == er.value,
           "abs_diff_oracle": abs(e.value - e_ref), "abs_diff_df": abs(ed.value - e.value),
           "e_df_panels_comm": ep.value, "abs_diff_df_panels": abs(ep.value - e.value),
           "panel_units_this_rank": int(sp.units), "panel_units_all_ranks": int(units_panel[0]),
           "units_this_rank": int(st.units), "h2d_bytes_all_ranks": float(h2d[0]), "dense_input_bytes": dense_bytes}
    assert same and rec["bitwise_equal_single_gpu"], rec
    assert rec["abs_diff_oracle"] < 1e-10 and rec["abs_diff_df"] < 1e-12 and rec["abs_diff_df_panels"] < 1e-12, rec
    assert st.units in (n // world, n // world + 1), rec
    assert rec["h2d_bytes_all_ranks"] < 1.25 * dense_bytes + world * (1 << 16) + 24 * n * world, rec
    out[f"o{o}_v{v}"] = rec
lib.mpqc_t_comm_destroy(comm)
if rank == 0:
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"rank_mode_check_n{world}.json"), "w"), indent=1)
dist.destroy_process_group()
