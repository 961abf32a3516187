"""Tensor-dump format for replaying a real MPQC (T) without Libint.

The adapter (integration/ccsd_t_gpu_impl.h, ``gpu_t::dump_problem``, keyword ``gpu_dump_file``) writes exactly what it hands to the C ABI; this module
reads/writes the same file so a dump produced where MPQC is installed can be replayed here (tests, bench) and a
fixture produced here (``oracle/h2o_golden.py``) can be loaded by a C++ host.

Layout (little endian):  8 bytes magic ``MPQCT001``; int64 o, v, n_frozen, n_all (= n_frozen + o + v);
then float64 arrays, row-major: eps[n_all], t1[v][o], t2[v][v][o][o], g_abij[v][v][o][o], g_aijk[v][o][o][o],
g_abci[v][v][v][o]  -- the reference's post-permutation layouts (ccsd_t.h:2219,2233,2242).
"""
from __future__ import annotations

import numpy as np

from .ccsd_t import DenseCCSD, InputError

MAGIC = b"MPQCT001"


def save_problem(path, eps_all, n_frozen, t1, t2, g_abij, g_aijk, g_abci):
    v, o = t1.shape
    eps_all = np.ascontiguousarray(eps_all, dtype="<f8")
    if eps_all.size != n_frozen + o + v:
        raise InputError("eps_all must hold frozen + active occupied + virtual orbital energies")
    with open(path, "wb") as f:
        f.write(MAGIC)
        np.array([o, v, n_frozen, eps_all.size], dtype="<i8").tofile(f)
        eps_all.tofile(f)
        for name, arr, shape in (("t1", t1, (v, o)), ("t2", t2, (v, v, o, o)), ("g_abij", g_abij, (v, v, o, o)),
                                 ("g_aijk", g_aijk, (v, o, o, o)), ("g_abci", g_abci, (v, v, v, o))):
            if tuple(arr.shape) != shape:
                raise InputError(f"{name} has shape {tuple(arr.shape)}, expected {shape}")
            np.ascontiguousarray(arr, dtype="<f8").tofile(f)


def load_problem(path, e_ccsd: float = 0.0) -> DenseCCSD:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise InputError(f"{path} is not an MPQCT001 dump")
        o, v, n_frozen, n_all = (int(x) for x in np.fromfile(f, dtype="<i8", count=4))
        if n_all != n_frozen + o + v or o < 1 or v < 1:
            raise InputError(f"{path}: inconsistent header")

        def rd(*shape):
            n = int(np.prod(shape))
            a = np.fromfile(f, dtype="<f8", count=n)
            if a.size != n:
                raise InputError(f"{path}: truncated")
            return a.reshape(shape)

        eps = rd(n_all)
        t1, t2 = rd(v, o), rd(v, v, o, o)
        g_abij, g_aijk, g_abci = rd(v, v, o, o), rd(v, o, o, o), rd(v, v, v, o)
    return DenseCCSD(t1, t2, g_abij, g_aijk, g_abci, eps, n_frozen, e_ccsd)
