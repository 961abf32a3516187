"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/mpqc_t.h declares, fails loudly (no CPU fallback) without a GPU, and its host-only helpers
agree with the oracle's unit enumeration."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from mpqc_b200 import build as B
from mpqc_b200 import lib as L
from oracle import ccsd_t_oracle as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


@pytest.fixture(scope="module")
def lib():
    B.build()
    return L.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mpqc_t.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mpqc_t_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header():
    # plain-C layout: 2 int64 + 7 pointers; options and stats are fixed-size PODs
    assert C.sizeof(L.Problem) == 2 * 8 + 7 * 8
    assert C.sizeof(L.Options) == 4 + 4 + 8 + 4 + 4 + 3 * 8 + 3 * 4 + 5 * 4
    assert C.sizeof(L.Stats) == 8 * 8 + 4 * 8 + 8 * 4
    assert C.sizeof(L.UniqueId) == 128


def test_version_and_strerror(lib):
    assert b"sm_100a" in lib.mpqc_t_version()
    assert lib.mpqc_t_strerror(L.OK) == b"ok"
    assert b"no CPU fallback" in lib.mpqc_t_strerror(L.ERR_NO_DEVICE)


@pytest.mark.parametrize("o", [1, 2, 3, 7, 21])
def test_unit_enumeration_matches_oracle(lib, o):
    tr = oc.ijk_triple_list(o)
    assert lib.mpqc_t_triple_count(o) == len(tr)
    i, j, k = C.c_int32(), C.c_int32(), C.c_int32()
    for u, t in enumerate(tr):
        assert lib.mpqc_t_triple_of_unit(o, u, C.byref(i), C.byref(j), C.byref(k)) == L.OK
        assert (i.value, j.value, k.value) == t
    assert lib.mpqc_t_triple_of_unit(o, len(tr), C.byref(i), C.byref(j), C.byref(k)) == L.ERR_BAD_ARG


def test_unit_enumeration_is_arithmetic_at_the_size_limit(lib):
    # units are decoded in closed form (no O(o^3) list): spot-check group boundaries at the largest accepted o
    o = 4096
    n = lib.mpqc_t_triple_count(o)
    assert n == o * (o + 1) * (o + 2) // 6 - o
    i, j, k = C.c_int32(), C.c_int32(), C.c_int32()

    def tr(u):
        assert lib.mpqc_t_triple_of_unit(o, u, C.byref(i), C.byref(j), C.byref(k)) == L.OK
        return (i.value, j.value, k.value)
    assert tr(0) == (1, 0, 0) and tr(1) == (1, 1, 0) and tr(2) == (2, 0, 0)
    assert tr(n - 1) == (o - 1, o - 1, o - 2) and tr(n - 2) == (o - 1, o - 1, o - 3)
    for a in (2, 63, 1000, 4095):
        first = a * (a + 1) * (a + 2) // 6 - a          # units before leading index a
        assert tr(first) == (a, 0, 0) and tr(first - 1) == (a - 1, a - 1, a - 2)
        b = a // 2
        assert tr(first + b * (b + 1) // 2 + b) == (a, b, b) and tr(first + b * (b + 1) // 2 + b + 1) == (a, b + 1, 0)


def test_flop_model(lib):
    assert lib.mpqc_t_flops(63, 297) == pytest.approx(oc.flops(63, 297))
    # per-unit count times the number of units approaches the 2 o^3 v^3 (v+o) model
    o, v = 63, 297
    tot = lib.mpqc_t_unit_flops(o, v) * lib.mpqc_t_triple_count(o)
    assert tot == pytest.approx(oc.flops(o, v), rel=0.06)


def test_bad_arguments_are_reported(lib):
    e = C.c_double()
    assert lib.mpqc_t_energy(None, None, C.byref(e), None) == L.ERR_BAD_ARG
    assert b"NULL" in lib.mpqc_t_last_error()
    h = C.c_void_p()
    assert lib.mpqc_t_create(C.byref(h), 0, 5, 0) == L.ERR_BAD_ARG
    assert lib.mpqc_t_create(None, 2, 5, 0) == L.ERR_BAD_ARG
    assert lib.mpqc_t_destroy(None) == L.OK


@pytest.mark.skipif(HAS_GPU, reason="checks the loud failure on a box without a GPU")
def test_no_gpu_means_error_not_fallback(lib):
    from mpqc_b200.synthetic import make_problem
    from mpqc_b200.ccsd_t import CCSD_T, DenseCCSD, FeatureDisabled
    p = make_problem(2, 3)
    prob = L.make_problem(2, 3, p["eps_occ"], p["eps_vir"], p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
    e = C.c_double(123.0)
    assert lib.mpqc_t_energy(C.byref(prob), None, C.byref(e), None) == L.ERR_NO_DEVICE
    h = C.c_void_p()
    assert lib.mpqc_t_create(C.byref(h), 2, 3, 0) == L.ERR_NO_DEVICE
    with pytest.raises(FeatureDisabled):
        CCSD_T({"type": "CCSD(T)"}, ccsd=DenseCCSD.from_problem(p)).compute_ccsd_t()


def test_make_problem_validates_shapes():
    from mpqc_b200.synthetic import make_problem
    p = make_problem(2, 3)
    with pytest.raises(ValueError):
        L.make_problem(2, 3, p["eps_occ"], p["eps_vir"], p["t1"].T.copy(), p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"])
    with pytest.raises(ValueError):
        L.make_problem(2, 3, p["eps_occ"], p["eps_vir"], p["t1"].astype(np.float32), p["t2"], p["g_abij"],
                       p["g_aijk"], p["g_abci"])


def _build_c_host(tmp_path, name="c_host"):
    import subprocess
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "mpqc_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-o", exe, "-L", libdir, "-lmpqc_t_cuda",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def _h2o_dump(tmp_path):
    from mpqc_b200 import dump
    g = np.load(os.path.join(ROOT, "tests", "golden", "h2o_631g.npz"))
    path = str(tmp_path / "h2o.mpqct")
    dump.save_problem(path, g["eps"], int(g["n_frozen"]), g["t1"], g["t2"], g["g_abij"], g["g_aijk"], g["g_abci"])
    return path


@pytest.mark.skipif(HAS_GPU, reason="CPU-box behaviour of the plain-C host")
def test_plain_c_host_links_and_fails_loudly_without_gpu(lib, tmp_path):
    # the header is consumable from C99 and the library from a C program; without a device: exit code 2
    import subprocess
    exe = _build_c_host(tmp_path)
    res = subprocess.run([exe, _h2o_dump(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 2 and "no CPU fallback" in res.stderr


@pytest.mark.skipif(HAS_GPU, reason="CPU-box behaviour of the plain-C host")
def test_plain_c_communicator_host_links_and_fails_loudly_without_gpu(lib, tmp_path):
    # the communicator API (mpqc_t_comm_*, mpqc_t_energy_comm, mpqc_t_host_alloc) is consumable from C99 as well
    import subprocess
    exe = _build_c_host(tmp_path, "c_host_comm")
    res = subprocess.run([exe, _h2o_dump(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 2 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_plain_c_host_reproduces_reference_h2o(lib, tmp_path):
    import subprocess
    exe = _build_c_host(tmp_path)
    res = subprocess.run([exe, _h2o_dump(tmp_path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    e = float(res.stdout.split("E(T) =")[1].split()[0])
    assert abs(e - (-0.000868413807153793)) < 1e-11
    assert "(T) Energy:" in res.stdout


def test_tiling_plan_of_the_baseline_configs(lib):
    # host-only planning: the W-contraction tiling for BASELINE.json's shapes and its padding efficiency
    def plan(o, v, flat):
        info = L.PlanInfo()
        assert lib.mpqc_t_plan(o, v, flat, C.byref(info)) == L.OK
        return info
    t = plan(63, 297, 1)                    # uracil trimer: flat rows, 3 column tiles of 13 fragments, last one short
    assert (t.kp, t.row_tiles, t.col_tiles, t.nfrag, t.skip_last) == (360, 690, 3, 13, 1)
    assert t.energy_tile_sets == 38 * 39 * 40 // 6
    assert 0.975 < t.flop_efficiency < 0.98            # 297/304 columns, 88209/88320 rows
    assert abs(t.bytes_operands - (2 * 63 * 297 * 297 * 360 + 63 * 63 * 297 * 360 + 63 * 63 * 297 * 297) * 8) < 1
    p = plan(63, 297, 0)                    # patch mode of the same shape: odd row patch, ~4% more padding
    assert p.tp * p.tq <= 128 and p.tp % 2 == 1 and 0.93 < p.flop_efficiency < t.flop_efficiency
    for (o, v, lo) in [(21, 93, 0.90), (42, 198, 0.97), (40, 530, 0.94), (50, 500, 0.975), (4, 8, 0.2)]:
        info = plan(o, v, 1)
        assert info.nfrag * 8 * info.col_tiles - 8 * info.skip_last >= v
        assert info.kp % 8 == 0 and info.kp >= v + o and info.flop_efficiency > lo, (o, v, info.flop_efficiency)
    bad = L.PlanInfo()
    assert lib.mpqc_t_plan(0, 5, 1, C.byref(bad)) == L.ERR_BAD_ARG


def test_density_fitted_memory_model(lib):
    # host-only model of the density-fitted path: with a panel cache the v^3 o operand is never
    # resident, which moves the single-GPU memory ceiling
    def plan(o, v, naux, block, flat=0):
        info = L.DfPlanInfo()
        assert lib.mpqc_t_plan_df(o, v, naux, block, flat, C.byref(info)) == L.OK
        return info
    res = plan(40, 530, 1140, 0)                 # (H2O)10 / cc-pVTZ, everything resident: 52 GB of panels alone
    assert not res.panel_mode and res.npanel == 40 and 51e9 < res.bytes_panels < 53e9
    pc = plan(40, 530, 1140, 6)                  # panel cache, occupied block 6: 18 panels
    assert pc.panel_mode and pc.npanel == 18 and pc.block == 6
    assert pc.bytes_total < 60e9 and pc.build_flop_fraction < 0.02
    big = plan(40, 1000, 3500, 3)                # v = 1000 on ONE GPU: A alone would be 333 GB
    assert big.panel_mode and big.bytes_total < 170e9 and big.build_flop_fraction < 0.08
    assert plan(40, 1000, 3500, 0).bytes_total > 333e9
    assert lib.mpqc_t_plan_df(0, 5, 5, 0, 0, C.byref(L.DfPlanInfo())) == L.ERR_BAD_ARG
