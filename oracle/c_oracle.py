"""ctypes loader for oracle/liboracle_t.so (plain-C restatement; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_t.so")


def build():
    src = os.path.join(_HERE, "ccsd_t_ref.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_t.so"], stdout=subprocess.DEVNULL)
    return _LIB


def straight_c(t1, t2, g_abij, g_aijk, g_abci, eps_all, n_frozen=0, mode=0) -> float:
    lib = C.CDLL(build())
    dp = C.POINTER(C.c_double)
    lib.mpqc_oracle_straight.restype = C.c_double
    lib.mpqc_oracle_straight.argtypes = [C.c_int, C.c_int, C.c_int] + [dp] * 6 + [C.c_int]
    v, o = t1.shape
    arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (eps_all, t1, t2, g_abij, g_aijk, g_abci)]
    return lib.mpqc_oracle_straight(o, v, n_frozen, *[a.ctypes.data_as(dp) for a in arrs], mode)
