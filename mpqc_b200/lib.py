"""ctypes binding of include/mpqc_t.h (libmpqc_t_cuda.so).

The library is the product; this module only marshals pointers.  It fails loudly when the shared
object is missing -- there is no Python/CPU fallback for the (T) path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPQC_T_LIB selects an experiment build of the same library (A/B runs, build.py --out=...); never a fallback
LIB_PATH = os.environ.get("MPQC_T_LIB") or os.path.join(_HERE, "libmpqc_t_cuda.so")

OK, ERR_BAD_ARG, ERR_NO_DEVICE, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_INTERNAL = range(7)

c_double_p = C.POINTER(C.c_double)


class Problem(C.Structure):
    _fields_ = [("o", C.c_int64), ("v", C.c_int64),
                ("eps_occ", C.c_void_p), ("eps_vir", C.c_void_p), ("t1", C.c_void_p), ("t2", C.c_void_p),
                ("g_abij", C.c_void_p), ("g_aijk", C.c_void_p), ("g_abci", C.c_void_p)]


class DfProblem(C.Structure):
    _fields_ = [("o", C.c_int64), ("v", C.c_int64), ("naux", C.c_int64),
                ("eps_occ", C.c_void_p), ("eps_vir", C.c_void_p), ("t1", C.c_void_p), ("t2", C.c_void_p),
                ("x_ab", C.c_void_p), ("x_ij", C.c_void_p), ("x_ai", C.c_void_p)]


class PlanInfo(C.Structure):
    _fields_ = [("kp", C.c_int64), ("flat", C.c_int32), ("tp", C.c_int32), ("tq", C.c_int32), ("row_tiles", C.c_int32),
                ("col_tiles", C.c_int32), ("nfrag", C.c_int32), ("skip_last", C.c_int32), ("energy_tile_sets", C.c_int32),
                ("flop_efficiency", C.c_double), ("bytes_operands", C.c_double)]


class Options(C.Structure):
    _fields_ = [("ngpu", C.c_int32), ("device_ids", C.POINTER(C.c_int32)), ("verbose", C.c_int32),
                ("inputs_on_device", C.c_int32), ("unit_first", C.c_int64), ("unit_stride", C.c_int64),
                ("unit_count", C.c_int64), ("batch", C.c_int32), ("steal_chunk", C.c_int32),
                ("use_nccl", C.c_int32), ("df_block", C.c_int32), ("reserved", C.c_int32 * 4)]


class Stats(C.Structure):
    _fields_ = [("seconds_total", C.c_double), ("seconds_upload", C.c_double), ("seconds_relayout", C.c_double),
                ("seconds_compute", C.c_double), ("seconds_contract", C.c_double), ("seconds_energy", C.c_double),
                ("flops", C.c_double), ("flops_executed", C.c_double), ("units", C.c_int64),
                ("kernel_launches", C.c_int64), ("bytes_h2d", C.c_int64), ("bytes_d2h", C.c_int64),
                ("ngpu", C.c_int32), ("reserved", C.c_int32 * 7)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class DfPlanInfo(C.Structure):
    _fields_ = [("npanel", C.c_int32), ("block", C.c_int32), ("panel_mode", C.c_int32), ("flat", C.c_int32),
                ("bytes_panels", C.c_double), ("bytes_b", C.c_double), ("bytes_gv", C.c_double), ("bytes_t2", C.c_double),
                ("bytes_factors", C.c_double), ("bytes_w_workspace", C.c_double), ("bytes_total", C.c_double),
                ("build_flop_fraction", C.c_double)]


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class MpqcTError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str):
        super().__init__(f"{what}: status {status} ({detail})")
        self.status = status


#: every symbol include/mpqc_t.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "mpqc_t_energy", "mpqc_t_energy_df", "mpqc_t_create", "mpqc_t_upload", "mpqc_t_upload_df", "mpqc_t_run", "mpqc_t_debug_w", "mpqc_t_stream",
    "mpqc_t_comm_unique_id", "mpqc_t_comm_create_rank", "mpqc_t_comm_create_local", "mpqc_t_comm_size", "mpqc_t_comm_release_cache", "mpqc_t_comm_destroy",
    "mpqc_t_energy_comm", "mpqc_t_energy_df_comm", "mpqc_t_host_alloc", "mpqc_t_host_free", "mpqc_t_run_vblocks", "mpqc_t_run_comm", "mpqc_t_set_df_block", "mpqc_t_plan_df", "mpqc_t_query", "mpqc_t_w_batch", "mpqc_t_shard_plan",
    "mpqc_t_destroy", "mpqc_t_triple_count", "mpqc_t_triple_of_unit", "mpqc_t_flops", "mpqc_t_unit_flops",
    "mpqc_t_device_count", "mpqc_t_plan", "mpqc_t_version", "mpqc_t_strerror", "mpqc_t_last_error", "mpqc_t_microbench",
]

QUERY_PANEL_SLOTS, QUERY_PANEL_MODE, QUERY_FLAT, QUERY_PANELS_BUILT, QUERY_PANEL_BLOCK = range(5)

_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python -m mpqc_b200.build` "
            "(the (T) path is CUDA-only, there is no fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
    vp = C.c_void_p
    lib.mpqc_t_energy.argtypes = [C.POINTER(Problem), C.POINTER(Options), c_double_p, C.POINTER(Stats)]
    lib.mpqc_t_energy.restype = C.c_int
    lib.mpqc_t_energy_df.argtypes = [C.POINTER(DfProblem), C.POINTER(Options), c_double_p, C.POINTER(Stats)]
    lib.mpqc_t_energy_df.restype = C.c_int
    lib.mpqc_t_upload_df.argtypes = [vp, C.POINTER(DfProblem), C.c_int32, C.POINTER(Stats)]
    lib.mpqc_t_upload_df.restype = C.c_int
    lib.mpqc_t_create.argtypes = [C.POINTER(vp), C.c_int64, C.c_int64, C.c_int32]
    lib.mpqc_t_create.restype = C.c_int
    lib.mpqc_t_upload.argtypes = [vp, C.POINTER(Problem), C.c_int32, C.POINTER(Stats)]
    lib.mpqc_t_upload.restype = C.c_int
    lib.mpqc_t_run.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p,
                               C.POINTER(Stats)]
    lib.mpqc_t_run.restype = C.c_int
    lib.mpqc_t_run_vblocks.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p, c_double_p,
                                       C.POINTER(Stats)]
    lib.mpqc_t_run_vblocks.restype = C.c_int
    lib.mpqc_t_run_comm.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, c_double_p, c_double_p,
                                    C.POINTER(Stats)]
    lib.mpqc_t_run_comm.restype = C.c_int
    lib.mpqc_t_shard_plan.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int64)]
    lib.mpqc_t_shard_plan.restype = C.c_int64
    lib.mpqc_t_w_batch.argtypes = [vp, C.POINTER(C.c_int32), C.c_int64, vp, C.c_int32]
    lib.mpqc_t_w_batch.restype = C.c_int
    lib.mpqc_t_query.argtypes = [vp, C.c_int32, C.POINTER(C.c_int64)]
    lib.mpqc_t_query.restype = C.c_int
    lib.mpqc_t_set_df_block.argtypes = [vp, C.c_int32]
    lib.mpqc_t_set_df_block.restype = C.c_int
    lib.mpqc_t_plan_df.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(DfPlanInfo)]
    lib.mpqc_t_plan_df.restype = C.c_int
    lib.mpqc_t_comm_unique_id.argtypes = [C.POINTER(UniqueId)]
    lib.mpqc_t_comm_unique_id.restype = C.c_int
    lib.mpqc_t_comm_create_rank.argtypes = [C.POINTER(vp), C.c_int32, C.c_int32, C.POINTER(UniqueId), C.c_int32]
    lib.mpqc_t_comm_create_rank.restype = C.c_int
    lib.mpqc_t_comm_create_local.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(C.c_int32)]
    lib.mpqc_t_comm_create_local.restype = C.c_int
    lib.mpqc_t_comm_size.argtypes = [vp]
    lib.mpqc_t_comm_size.restype = C.c_int
    lib.mpqc_t_comm_release_cache.argtypes = [vp]
    lib.mpqc_t_comm_release_cache.restype = C.c_int
    lib.mpqc_t_comm_destroy.argtypes = [vp]
    lib.mpqc_t_comm_destroy.restype = C.c_int
    lib.mpqc_t_energy_comm.argtypes = [vp, C.POINTER(Problem), C.POINTER(Options), c_double_p, C.POINTER(Stats)]
    lib.mpqc_t_energy_comm.restype = C.c_int
    lib.mpqc_t_energy_df_comm.argtypes = [vp, C.POINTER(DfProblem), C.POINTER(Options), c_double_p, C.POINTER(Stats)]
    lib.mpqc_t_energy_df_comm.restype = C.c_int
    lib.mpqc_t_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.mpqc_t_host_alloc.restype = C.c_int
    lib.mpqc_t_host_free.argtypes = [vp]
    lib.mpqc_t_host_free.restype = C.c_int
    lib.mpqc_t_debug_w.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, c_double_p]
    lib.mpqc_t_debug_w.restype = C.c_int
    lib.mpqc_t_stream.argtypes = [vp]
    lib.mpqc_t_stream.restype = vp
    lib.mpqc_t_destroy.argtypes = [vp]
    lib.mpqc_t_destroy.restype = C.c_int
    lib.mpqc_t_triple_count.argtypes = [C.c_int64]
    lib.mpqc_t_triple_count.restype = C.c_int64
    lib.mpqc_t_triple_of_unit.argtypes = [C.c_int64, C.c_int64] + [C.POINTER(C.c_int32)] * 3
    lib.mpqc_t_triple_of_unit.restype = C.c_int
    lib.mpqc_t_flops.argtypes = [C.c_int64, C.c_int64]
    lib.mpqc_t_flops.restype = C.c_double
    lib.mpqc_t_unit_flops.argtypes = [C.c_int64, C.c_int64]
    lib.mpqc_t_unit_flops.restype = C.c_double
    lib.mpqc_t_plan.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.POINTER(PlanInfo)]
    lib.mpqc_t_plan.restype = C.c_int
    lib.mpqc_t_device_count.argtypes = []
    lib.mpqc_t_device_count.restype = C.c_int
    lib.mpqc_t_version.restype = C.c_char_p
    lib.mpqc_t_strerror.argtypes = [C.c_int]
    lib.mpqc_t_strerror.restype = C.c_char_p
    lib.mpqc_t_last_error.restype = C.c_char_p
    lib.mpqc_t_microbench.argtypes = [C.c_int32, C.c_int32, c_double_p]
    lib.mpqc_t_microbench.restype = C.c_int
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != OK:
        lib = load()
        detail = lib.mpqc_t_strerror(status).decode() + "; " + lib.mpqc_t_last_error().decode()
        raise MpqcTError(status, what, detail)


def _ptr(x) -> int:
    """Address of a numpy array (host) or torch tensor (host or device)."""
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64 or not x.flags["C_CONTIGUOUS"]:
            raise ValueError("tensors must be C-contiguous float64")
        return x.ctypes.data
    # torch tensor
    import torch
    if x.dtype != torch.float64 or not x.is_contiguous():
        raise ValueError("tensors must be contiguous float64")
    return x.data_ptr()


def make_problem(o, v, eps_occ, eps_vir, t1, t2, g_abij, g_aijk, g_abci):
    """Build the C struct; the caller must keep the arrays alive while it is in use."""
    shapes = dict(eps_occ=(o,), eps_vir=(v,), t1=(v, o), t2=(v, v, o, o), g_abij=(v, v, o, o),
                  g_aijk=(v, o, o, o), g_abci=(v, v, v, o))
    arrs = dict(eps_occ=eps_occ, eps_vir=eps_vir, t1=t1, t2=t2, g_abij=g_abij, g_aijk=g_aijk, g_abci=g_abci)
    for k, shp in shapes.items():
        if tuple(arrs[k].shape) != shp:
            raise ValueError(f"{k} has shape {tuple(arrs[k].shape)}, expected {shp}")
    p = Problem(o=o, v=v, **{k: _ptr(a) for k, a in arrs.items()})
    p._keepalive = arrs
    return p


def make_df_problem(o, v, naux, eps_occ, eps_vir, t1, t2, x_ab, x_ij, x_ai):
    """Density-fitted inputs: x_ab[naux,v,v] (symmetric in a,b), x_ij[naux,o,o], x_ai[naux,v,o]."""
    shapes = dict(eps_occ=(o,), eps_vir=(v,), t1=(v, o), t2=(v, v, o, o), x_ab=(naux, v, v), x_ij=(naux, o, o),
                  x_ai=(naux, v, o))
    arrs = dict(eps_occ=eps_occ, eps_vir=eps_vir, t1=t1, t2=t2, x_ab=x_ab, x_ij=x_ij, x_ai=x_ai)
    for k, shp in shapes.items():
        if tuple(arrs[k].shape) != shp:
            raise ValueError(f"{k} has shape {tuple(arrs[k].shape)}, expected {shp}")
    p = DfProblem(o=o, v=v, naux=naux, **{k: _ptr(a) for k, a in arrs.items()})
    p._keepalive = arrs
    return p
