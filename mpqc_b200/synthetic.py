"""Synthetic (T) inputs with the true permutational symmetries.

ERIs come from a random DF-like factor L[P,p,q] = L[P,q,p], (pq|rs) = s * sum_P L[P,pq] L[P,rs],
sliced with the index mappings of the reference's integral getters
(/root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:2219,2233,2242):

    g_abij[a,b,i,j] = <ij|ab> = (ia|jb)
    g_aijk[a,i,j,k] = <ij|ka> = (ik|ja)
    g_abci[a,b,c,i] = <ia|bc> = (ib|ac)

T2 is MP1-like (g_abij / D2, so t2[a,b,i,j] == t2[b,a,j,i]); T1 ~ N(0, 0.02^2).
Orbital energies are sorted uniform draws with a gap, so every triples denominator has
|D| >= 1.5.  ``numpy`` version for the CPU tests, ``torch`` version for HBM-resident inputs.
"""
from __future__ import annotations

import numpy as np

SEED = 20261017


def calibrated_scale(o: int, v: int, target: float = 0.3) -> float:
    """ERI scale that lands |E(T)| of the synthetic problem near ``target`` Eh (within ~10x), so the
    1e-9 Eh absolute parity tolerance is a ~1e-9 relative one.  E(T) grows like scale^4; the
    size dependence E(1) ~ 38 (o^3 v^3 / 3.3e4)^0.75 is an empirical fit over o=4..63, v=8..297."""
    e1 = 38.0 * (float(o) ** 3 * float(v) ** 3 / 3.3e4) ** 0.75
    return float((target / e1) ** 0.25)


def _eps(rng, o, v):
    eps_occ = np.sort(rng.uniform(-1.5, -0.3, size=o))
    eps_vir = np.sort(rng.uniform(0.2, 3.0, size=v))
    return eps_occ, eps_vir


def make_problem(o: int, v: int, seed: int = SEED, scale: float | None = None, naux: int | None = None):
    """Return dict of numpy float64 arrays in the reference layouts."""
    if scale is None:
        scale = calibrated_scale(o, v)
    rng = np.random.default_rng(seed)
    naux = naux or 2 * (o + v)
    eps_occ, eps_vir = _eps(rng, o, v)
    s = np.sqrt(scale / naux)
    l_oo = rng.standard_normal((naux, o, o)) * s
    l_oo = 0.5 * (l_oo + l_oo.transpose(0, 2, 1))
    l_vv = rng.standard_normal((naux, v, v)) * s
    l_vv = 0.5 * (l_vv + l_vv.transpose(0, 2, 1))
    l_ov = rng.standard_normal((naux, o, v)) * s
    g_abij = np.einsum("Pia,Pjb->abij", l_ov, l_ov, optimize=True)
    g_aijk = np.einsum("Pik,Pja->aijk", l_oo, l_ov, optimize=True)
    g_abci = np.einsum("Pib,Pac->abci", l_ov, l_vv, optimize=True)
    d2 = (eps_occ[None, None, :, None] + eps_occ[None, None, None, :]
          - eps_vir[:, None, None, None] - eps_vir[None, :, None, None])
    t2 = g_abij / d2
    t1 = rng.standard_normal((v, o)) * 0.02
    return dict(o=o, v=v, eps_occ=eps_occ, eps_vir=eps_vir,
                t1=np.ascontiguousarray(t1), t2=np.ascontiguousarray(t2),
                g_abij=np.ascontiguousarray(g_abij), g_aijk=np.ascontiguousarray(g_aijk),
                g_abci=np.ascontiguousarray(g_abci),
                # the three-centre factors in the reference's layouts (CCSD::get_Xab/Xij/Xai, ccsd.h:480-493)
                naux=naux, x_ab=np.ascontiguousarray(l_vv), x_ij=np.ascontiguousarray(l_oo),
                x_ai=np.ascontiguousarray(l_ov.transpose(0, 2, 1)))


def make_problem_torch(o: int, v: int, device, seed: int = SEED, scale: float | None = None,
                       naux: int | None = None, dense_abci: bool = True):
    """Same construction on a CUDA device with torch (plumbing: used only to put synthetic
    inputs into HBM for bench.py / large-size tests; not bit-identical to the numpy version).
    ``dense_abci=False`` skips the v^3 o tensor (density-fitted hand-off: only the three-centre factors
    exist); every other array is identical to the dense call with the same seed."""
    import torch
    if scale is None:
        scale = calibrated_scale(o, v)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    naux = naux or 2 * (o + v)
    eps_occ = torch.sort(torch.rand(o, generator=g, device=device, dtype=f64) * 1.2 - 1.5).values
    eps_vir = torch.sort(torch.rand(v, generator=g, device=device, dtype=f64) * 2.8 + 0.2).values
    s = (scale / naux) ** 0.5
    l_oo = torch.randn(naux, o, o, generator=g, device=device, dtype=f64) * s
    l_oo = 0.5 * (l_oo + l_oo.transpose(1, 2))
    l_vv = torch.randn(naux, v, v, generator=g, device=device, dtype=f64) * s
    l_vv = 0.5 * (l_vv + l_vv.transpose(1, 2))
    l_ov = torch.randn(naux, o, v, generator=g, device=device, dtype=f64) * s
    lov2 = l_ov.reshape(naux, o * v)
    # (ia|jb) -> [a,b,i,j]
    g_iajb = (lov2.t() @ lov2).reshape(o, v, o, v)
    g_abij = g_iajb.permute(1, 3, 0, 2).contiguous()
    del g_iajb
    # (ik|ja) -> [a,i,j,k]
    g_ikja = (l_oo.reshape(naux, o * o).t() @ lov2).reshape(o, o, o, v)
    g_aijk = g_ikja.permute(3, 0, 2, 1).contiguous()
    del g_ikja
    # (ib|ac) -> [a,b,c,i]; build per a-slab to bound the transient
    g_abci = None
    if dense_abci:
        g_abci = torch.empty(v, v, v, o, device=device, dtype=f64)
        lvv2 = l_vv.reshape(naux, v * v)
        slab = max(1, min(v, (1 << 28) // max(1, v * v * o)))
        for a0 in range(0, v, slab):
            a1 = min(v, a0 + slab)
            # (ac|ib): rows (a,c) for a in slab
            blk = lvv2[:, a0 * v:a1 * v].t() @ lov2                 # [(a c), (i b)]
            blk = blk.reshape(a1 - a0, v, o, v)                      # a c i b
            g_abci[a0:a1] = blk.permute(0, 3, 1, 2)                  # a b c i
            del blk
    d2 = (eps_occ[None, None, :, None] + eps_occ[None, None, None, :]
          - eps_vir[:, None, None, None] - eps_vir[None, :, None, None])
    t2 = (g_abij / d2).contiguous()
    t1 = torch.randn(v, o, generator=g, device=device, dtype=f64) * 0.02
    return dict(o=o, v=v, eps_occ=eps_occ, eps_vir=eps_vir, t1=t1, t2=t2,
                g_abij=g_abij, g_aijk=g_aijk, g_abci=g_abci, naux=naux, x_ab=l_vv.contiguous(),
                x_ij=l_oo.contiguous(), x_ai=l_ov.transpose(1, 2).contiguous())
