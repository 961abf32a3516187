"""Generates tests/golden/synthetic_*.npz: seeded synthetic inputs (regenerable from the seed, so only
o, v, seed are stored) with E(T) from the three oracle restatements and per-unit energies.

The reference is C++ on TiledArray/MADWorld and cannot be built or imported in this container
(SURVEY.md section 8c), so these goldens are ORACLE-generated: they pin the oracle against itself across
time (regression) and pin the CUDA path against the oracle; they are not reference-generated.
The reference-generated pin is tests/golden/h2o_631g_*.npz (see oracle/h2o_golden.py).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mpqc_b200.synthetic import make_problem, SEED  # noqa: E402
from oracle import ccsd_t_oracle as oc  # noqa: E402

CASES = [(2, 3), (3, 5), (4, 8), (5, 19), (4, 24), (7, 17), (6, 33)]

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for idx, (o, v) in enumerate(CASES):
        seed = SEED + idx
        p = make_problem(o, v, seed=seed)
        args = (p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], p["eps_occ"], p["eps_vir"])
        e_ijk, parts = oc.ijk_driven(*args, return_parts=True)
        e_coarse = oc.coarse(*args, vir_block=8)
        e_straight = oc.straight(*args) if o ** 3 * v ** 3 <= 4e6 else np.nan
        np.savez(os.path.join(here, f"synthetic_o{o}_v{v}.npz"), o=o, v=v, seed=seed, e_ijk=e_ijk,
                 e_coarse=e_coarse, e_straight=e_straight, unit_e=parts,
                 checksum=np.array([p[k].sum() for k in ("t1", "t2", "g_abij", "g_aijk", "g_abci")]))
        print(o, v, e_ijk, e_coarse, e_straight)
