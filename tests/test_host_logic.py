"""CPU tests of the host-side mirror of the reference interface and of the multi-rank sharding
(world_size-2 gloo): KeyVal keywords, error behaviour, unit sharding + final sum."""
import io
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpqc_b200.ccsd_t import (CCSD_T, CCSD_T_F12, DenseCCSD, Energy, FeatureDisabled, InputError, TRange1Engine, class_ptr)
from mpqc_b200.synthetic import make_problem
from oracle import ccsd_t_oracle as oc


def test_keyval_keywords_and_defaults():
    w = CCSD_T({"type": "CCSD(T)"})
    assert w.approach_ == "gpu" and w.occ_block_size_ == 8 and w.unocc_block_size_ == 8
    assert not w.reblock_ and not w.reblock_inner_ and w.increase_ == 2 and w.n_laplace_quad_ == 4
    # the reference's golden input (tests/validation/reference/inputs/h2o-ccsd_t-631g-pvdz.json:34-38)
    w = CCSD_T({"type": "CCSD(T)", "method": "df", "approach": "straight", "occ_block_size": 4,
                "unocc_block_size": 4, "reblock_occ": 4, "reblock_unocc": 4})
    assert w.approach_ == "straight" and w.reblock_ and w.occ_block_size_ == 4
    w = CCSD_T({"approach": "laplace", "reblock_occ": 4, "reblock_inner": 7, "quadrature_points": 3})
    assert not w.reblock_ and not w.reblock_inner_ and w.n_laplace_quad_ == 3      # ccsd_t.h:125-128


def test_gpu_keywords_follow_the_patched_reference():
    # integration/mpqc_ccsd_t_gpu.patch adds ngpu / gpu_batch / gpu_df / gpu_dump_file to the reference's KeyVal table
    w = CCSD_T({"type": "CCSD(T)", "ngpu": 2, "gpu_batch": 3, "gpu_df": True, "gpu_df_block": 4, "gpu_dump_file": "x.mpqct"})
    assert (w.ngpu_, w.batch_, w.df_direct_, w.df_block_, w.dump_file_) == (2, 3, True, 4, "x.mpqct")
    w = CCSD_T({"batch": 5, "df_direct": True})                 # round-1 spellings keep working
    assert w.batch_ == 5 and w.df_direct_ and w.df_block_ == 0 and w.dump_file_ == ""
    patch = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "integration",
                              "mpqc_ccsd_t_gpu.patch")).read()
    for key in ("ngpu", "gpu_batch", "gpu_df", "gpu_dump_file"):
        assert f'"{key}"' in patch, key


def test_invalid_approach_raises_input_error():
    with pytest.raises(InputError) as ei:
        CCSD_T({"type": "CCSD(T)", "approach": "medium"})
    assert ei.value.keyword == "approach" and "Invalid (T) approach" in str(ei.value)     # ccsd_t.h:118-122
    with pytest.raises(InputError):
        class_ptr({"type": "CCSD(Q)"})
    with pytest.raises(InputError):
        CCSD_T({"rank": 2, "world_size": 2})
    assert isinstance(class_ptr({"type": "CCSD(T)"}), CCSD_T)


def test_second_caller_is_registered_and_shares_the_dispatcher():
    # "CCSD(T)F12" (f12/ccsd_t_f12.h:47-69) is the path's second caller: same keywords, same compute_ccsd_t()
    w = class_ptr({"type": "CCSD(T)F12", "approach": "coarse", "reblock_occ": 4})
    assert isinstance(w, CCSD_T_F12) and isinstance(w, CCSD_T) and w.approach_ == "coarse" and w.reblock_
    assert CCSD_T_F12.compute_ccsd_t is CCSD_T.compute_ccsd_t          # not overridden: the base's dispatcher runs
    with pytest.raises(InputError):
        CCSD_T_F12({"type": "CCSD(T)"})


def test_laplace_is_feature_disabled():
    p = make_problem(2, 3)
    w = CCSD_T({"approach": "laplace"}, ccsd=DenseCCSD.from_problem(p), out=io.StringIO())
    with pytest.raises(FeatureDisabled):
        w.compute_ccsd_t()


def test_trange1_engine_and_frozen_core_slicing():
    eng = TRange1Engine(n_occ=5, n_all=13, n_frozen=1)
    assert (eng.get_occ(), eng.get_nfrozen(), eng.get_active_occ(), eng.get_vir()) == (5, 1, 4, 8)
    p = make_problem(4, 8)
    cc = DenseCCSD.from_problem(p, n_frozen=1)
    eps = cc.orbital_energy()
    assert len(eps) == 13
    np.testing.assert_array_equal(eps[1:5], p["eps_occ"])       # eps[i + n_frozen]  (ccsd_t.h:2306-2311)
    np.testing.assert_array_equal(eps[5:], p["eps_vir"])        # eps[a + n_occ]
    with pytest.raises(InputError):
        DenseCCSD(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], eps[:-1], 1)


def test_obsolete_resets_state():
    w = CCSD_T({})
    w.triples_energy_, w.computed_ = -1.0, True
    w.obsolete()
    assert w.triples_energy() == 0.0 and not w.computed()
    assert w.can_evaluate(Energy()) and not w.can_evaluate(object())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, o, v, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_problem(o, v)
    args = (p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    # the shard this rank's CCSD_T would hand to mpqc_t_energy: unit_first=rank, unit_stride=world
    def reduce(x):                                      # replaces gop.sum (ccsd_t.h:692); NCCL on the GPU box
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])
    w = CCSD_T({"rank": rank, "world_size": world}, reduce=reduce)
    with pytest.raises(InputError):                     # a sharded run without a reducer never stores a partial energy
        CCSD_T({"rank": rank, "world_size": world}, ccsd=object()).compute_ccsd_t()
    units = oc.ijk_triple_list(o)[w.rank_::w.world_size_]
    partial = oc.ijk_driven(*args, triples=units)       # stand-in for the device result of that shard
    t = torch.tensor([partial], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)            # replaces gop.sum (ccsd_t.h:692); NCCL on the GPU box
    q.put((rank, len(units), float(t[0])))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    o, v = 4, 6
    p = make_problem(o, v)
    ref = oc.coarse(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, o, v, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    n_units = sum(r[1] for r in res)
    assert n_units == o * (o + 1) * (o + 2) // 6 - o
    for r in res:
        assert abs(r[2] - ref) < 1e-12


def _shard(lib, o, nranks, rank, all_local, panel_block, first=0, stride=1, count=-1):
    import ctypes as C
    n = lib.mpqc_t_triple_count(o)
    buf = (C.c_int64 * n)()
    tail = C.c_int64()
    m = lib.mpqc_t_shard_plan(o, first, stride, count, nranks, rank, all_local, panel_block, buf, n, C.byref(tail))
    return list(buf[:m]), tail.value


def _plan_rank_main(rank, world, port, o, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpqc_b200 import lib as L
    lib = L.load()
    p = make_problem(o, 5)
    args = (p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    units = oc.ijk_triple_list(o)
    out = []
    for panel_block in (0, 2):          # unit-cyclic, and occupied-block groups (panel-cache mode)
        mine, _ = _shard(lib, o, world, rank, 0, panel_block)           # rank mode: the library's own split
        partial = oc.ijk_driven(*args, triples=[units[u] for u in mine])
        t = torch.tensor([partial], dtype=torch.float64)
        dist.all_reduce(t)                                              # what ncclAllReduce does on the GPU box
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        out.append((panel_block, float(t[0]), gathered))
    q.put((rank, out))
    dist.destroy_process_group()


def test_library_shard_plan_two_ranks_gloo():
    # the split mpqc_t_energy_comm performs inside (host-only export mpqc_t_shard_plan), driven by two gloo ranks: the
    # shards partition the job in both modes and the reduced energy is the oracle's
    o = 6
    p = make_problem(o, 5)
    ref = oc.ijk_driven(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
    n = o * (o + 1) * (o + 2) // 6 - o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plan_rank_main, args=(r, 2, port, o, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
    for r in range(2):
        for panel_block, e, gathered in res[r]:
            assert abs(e - ref) < 1e-12
            assert sorted(gathered[0] + gathered[1]) == list(range(n))          # a partition of the job
            if panel_block == 0:
                assert gathered[0] == list(range(0, n, 2)) and gathered[1] == list(range(1, n, 2))
            else:
                assert abs(len(gathered[0]) - len(gathered[1])) <= 8            # groups of <= 8 units, dealt greedily


def test_library_shard_plan_properties():
    from mpqc_b200 import lib as L
    lib = L.load()
    o, n = 21, 1750
    for W in (1, 2, 4, 8):
        # one process driving W GPUs: static 7/8 (a multiple of W), the rest is the work-stealing tail
        shards = [_shard(lib, o, W, r, 1, 0) for r in range(W)]
        tail = shards[0][1]
        assert tail == (n if W == 1 else (n // 8) * 7 // W * W) and all(t == tail for _, t in shards)
        assert sorted(sum((s for s, _ in shards), [])) == list(range(tail))
        # panel-cache mode: whole occupied-block groups per worker, no tail, balanced within one group
        ps = [_shard(lib, o, W, r, 1, 4)[0] for r in range(W)]
        assert sorted(sum(ps, [])) == list(range(n))
        assert max(len(s) for s in ps) - min(len(s) for s in ps) <= 64
        i, j, k = (__import__("ctypes").c_int32() for _ in range(3))
        keys_of = []
        for s in ps:
            ks = set()
            for u in s:
                lib.mpqc_t_triple_of_unit(o, u, i, j, k)
                ks.add((i.value // 4, j.value // 4, k.value // 4))
            keys_of.append(ks)
        for a in range(W):
            for b in range(a + 1, W):
                assert not (keys_of[a] & keys_of[b])                             # a group never straddles two workers
    # a sub-job (first, stride, count) and bad arguments
    mine, _ = _shard(lib, o, 3, 1, 0, 0, first=5, stride=2, count=100)
    assert mine == list(range(1, 100, 3))
    assert lib.mpqc_t_shard_plan(o, 0, 1, -1, 2, 2, 0, 0, None, 0, None) == -1


def test_dump_roundtrip_and_h2o_fixture(tmp_path):
    from mpqc_b200 import dump
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "h2o_631g.npz"))
    path = str(tmp_path / "h2o.mpqct")
    dump.save_problem(path, g["eps"], int(g["n_frozen"]), g["t1"], g["t2"], g["g_abij"], g["g_aijk"], g["g_abci"])
    cc = dump.load_problem(path, e_ccsd=-1.0)
    eng = cc.trange1_engine()
    assert (eng.get_active_occ(), eng.get_vir(), eng.get_nfrozen()) == (4, 8, 1)
    np.testing.assert_array_equal(cc.get_abci(), g["g_abci"])
    np.testing.assert_array_equal(cc.orbital_energy(), g["eps"])
    e = oc.ijk_driven(cc.t1(), cc.t2(), cc.get_abij(), cc.get_aijk(), cc.get_abci(), cc.orbital_energy()[1:5],
                      cc.orbital_energy()[5:])
    assert abs(e - (-0.000868413807153793)) < 1e-11
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(InputError):
        dump.load_problem(path)


def test_bench_reference_arm_json_contract():
    # the CPU arm of bench.py on a tiny shape: one JSON line with the keys the driver reads
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload",
                          "tiny-selftest", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env,
                         timeout=300)
    assert res.returncode == 0, res.stderr
    line = json.loads([x for x in res.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
