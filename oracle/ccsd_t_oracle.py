"""CPU oracle for MPQC's closed-shell CCSD (T) energy correction.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mpqc_b200/`` may import this module: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` use it, and only as the checker / the CPU arm, never as the product path.

PARITY PINNING.  The reference holds no test that pins E(T) at the (tensors -> E(T)) boundary (SURVEY.md
section 8c) and cannot be compiled here (TiledArray / MADNESS / Libint2 / Boost / TBB / Eigen / MPI absent); its only
pin is the end-to-end validation case ``tests/validation/reference/outputs/h2o-ccsd_t-631g-pvdz.out:395``,
(T) = -0.000868413807153793 Eh.  ``oracle/h2o_golden.py`` rebuilds that case from scratch (own Gaussian
integrals, DF-RHF, spin-orbital DF-CCSD), reproduces the stored SCF (1e-12) and MP2 (2e-12) energies, and feeds the
resulting tensors to the functions below: all three restatements return the reference's stored (T) to 4e-13 Eh
(``tests/golden/h2o_631g.npz``, ``tests/test_oracle.py::test_h2o_reference_golden``).  STATUS: PINNED against a
reference-produced number.  In addition the three restatements (straight, coarse, ijk-driven) are pinned against
each other and against the plain-C literal restatement in ``oracle/ccsd_t_ref.c``.

All citations are ``file:line`` in ``/root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h``
unless another file is named.

Tensor layouts are the reference's post-permutation layouts (row-major, last index fastest):
    t1[a,i]                      ccsd.h:165-179
    t2[a,b,i,j]                  ccsd.h:165-179
    g_abij[a,b,i,j] = <ij|ab>    ccsd_t.h:2238-2244
    g_aijk[a,i,j,k] = <ij|ka>    ccsd_t.h:2210-2221   (consumed as g_cjkl[c,j,k,l])
    g_abci[a,b,c,i] = <ia|bc>    ccsd_t.h:2224-2235   (consumed as g_dabi[d,a,b,i])
    eps_occ[i] = eps[n_frozen + i], eps_vir[a] = eps[n_occ + a]   ccsd_t.h:2299-2311
"""
from __future__ import annotations

import itertools
import numpy as np

__all__ = [
    "straight", "coarse", "ijk_driven", "ijk_triple_list", "triple_weight",
    "w_ijk", "v_ijk", "energy_ijk", "reduce_plain", "reduce_symm", "flops", "vblock_energies",
]


def flops(o: int, v: int) -> float:
    """Algorithmic FLOP count 2 o^3 v^3 (v+o)."""
    return 2.0 * o ** 3 * v ** 3 * (v + o)


# ----------------------------------------------------------------------------------------
# Oracle A: compute_ccsd_t_straight, ccsd_t.h:1127-1170 (literal einsum transcription)
# ----------------------------------------------------------------------------------------
def _tr(x, spec_out: str, spec_in: str = "abcijk"):
    """TiledArray-style annotation permute: result(spec_in) = x(spec_out).

    ``r("a,b,c,i,j,k") = x("a,c,b,i,k,j")`` means r[a,b,c,i,j,k] = x[a,c,b,i,k,j].
    """
    # x axes are labelled by spec_out; we want axes ordered as spec_in.
    return np.einsum(f"{spec_out}->{spec_in}", x)


def straight(t1, t2, g_abij, g_aijk, g_abci, eps_occ, eps_vir) -> float:
    """E(T) by the reference's 'straight' approach, ccsd_t.h:1127-1170.  O(o^3 v^3) memory."""
    x = np.einsum("dabi,dcjk->abcijk", g_abci, t2, optimize=True)
    x = x - np.einsum("cjkl,abil->abcijk", g_aijk, t2, optimize=True)
    # :1144-1146
    t3 = (x + _tr(x, "acbikj") + _tr(x, "cabkij") + _tr(x, "cbakji")
          + _tr(x, "bcajki") + _tr(x, "bacjik"))
    # :1150-1152
    y = np.einsum("abij,ck->abcijk", g_abij, t1)
    v3 = y + _tr(y, "bcajki") + _tr(y, "acbikj")
    # :1163-1167
    z = (4.0 * t3 + _tr(t3, "abckij") + _tr(t3, "abcjki")
         - 2.0 * (_tr(t3, "abckji") + _tr(t3, "abcikj") + _tr(t3, "abcjik")))
    result = (t3 + v3) * z
    o, v = len(eps_occ), len(eps_vir)
    off = (0, 0, 0, 0, 0, 0)
    e = reduce_plain(result, eps_occ, eps_vir, off)
    return e / 3.0                                   # :1168


# ----------------------------------------------------------------------------------------
# The two reducers (MPQC's own arithmetic on the path), ccsd_t.h:2274-2432
# ----------------------------------------------------------------------------------------
def _denominator(eps_occ, eps_vir, a_idx, b_idx, c_idx):
    ea = eps_vir[a_idx][:, None, None, None, None, None]
    eb = eps_vir[b_idx][None, :, None, None, None, None]
    ec = eps_vir[c_idx][None, None, :, None, None, None]
    ei = eps_occ[:, None, None][None, None, None]
    ej = eps_occ[None, :, None][None, None, None]
    ek = eps_occ[None, None, :][None, None, None]
    # :2306-2320   e_abcijk = eps_i + eps_j + eps_k - eps_a - eps_b - eps_c
    return ei + ej + ek - ea - eb - ec


def reduce_plain(tile, eps_occ, eps_vir, offset) -> float:
    """CCSD_T_Reduce::operator(), ccsd_t.h:2286-2334: sum tile/(e_i+e_j+e_k-e_a-e_b-e_c)."""
    na, nb, nc = tile.shape[:3]
    a_idx = np.arange(na) + offset[0]
    b_idx = np.arange(nb) + offset[1]
    c_idx = np.arange(nc) + offset[2]
    d = _denominator(eps_occ, eps_vir, a_idx, b_idx, c_idx)
    return float(np.sum(tile / d))


def reduce_symm(tile, eps_occ, eps_vir, offset) -> float:
    """CCSD_T_ReduceSymm::operator(), ccsd_t.h:2350-2431.

    Only global c <= b <= a contribute; weight 2 if a,b,c all distinct, 0 if a==b==c,
    1 otherwise (:2399-2423).
    """
    na, nb, nc = tile.shape[:3]
    a_idx = np.arange(na) + offset[0]
    b_idx = np.arange(nb) + offset[1]
    c_idx = np.arange(nc) + offset[2]
    d = _denominator(eps_occ, eps_vir, a_idx, b_idx, c_idx)
    A = a_idx[:, None, None]
    B = b_idx[None, :, None]
    C = c_idx[None, None, :]
    keep = (B <= A) & (C <= B)
    none_equal = (A != B) & (A != C) & (B != C)
    diagonal = (A == B) & (B == C)
    w = np.where(none_equal, 2.0, np.where(diagonal, 0.0, 1.0)) * keep
    return float(np.sum((tile / d) * w[:, :, :, None, None, None]))


# ----------------------------------------------------------------------------------------
# Oracle B: compute_ccsd_t_coarse_grain, ccsd_t.h:200-711 (the reference's default approach)
# ----------------------------------------------------------------------------------------
def _blocks(n: int, bs: int):
    """TRange1Engine::compute_trange1-style tiling (expression/trange1_engine.cpp:10-20)."""
    nb = max(1, (n + bs - 1) // bs)
    # the reference spreads the remainder; any tiling is numerically the identity
    edges = [min(n, i * bs) for i in range(nb)] + [n]
    return [(edges[i], edges[i + 1]) for i in range(nb) if edges[i + 1] > edges[i]]


def coarse(t1, t2, g_abij, g_aijk, g_abci, eps_occ, eps_vir, vir_block: int = 8,
           rank: int = 0, size: int = 1, block_filter=None, return_parts: bool = False):
    """E(T) by the reference's default 'coarse' approach, ccsd_t.h:443-640.

    Loops a >= b >= c over virtual *blocks* holding all occupied i,j,k; each of the six
    contraction pairs is a matmul (the reference's TA contraction -> BLAS dgemm, :314-340,
    :359-374), followed by the 6-index permutes (:498-555), the V outer products (:559-596),
    the symmetrise-multiply (:598-609) and the reducers (:611-640).
    ``rank/size`` reproduce the round-robin work split (:477-480).  ``block_filter`` (a set of
    global_iter values) restricts the loop for sampled CPU-baseline timing.
    """
    o, v = len(eps_occ), len(eps_vir)
    vb = _blocks(v, vir_block)
    t2_abil = np.ascontiguousarray(t2)                       # [a,b,i,l]
    g_dabi = g_abci                                          # [d,a,b,i]
    e_total = 0.0
    parts = []
    global_iter = 0

    def t3_term(a, b, c):
        """block (g_dabi * t2_dcjk) - (t2_abil * g_cjkl) -> t3[a,b,i,c,j,k]   (:314-340)"""
        (a0, a1), (b0, b1), (c0, c1) = a, b, c
        na, nb_, nc = a1 - a0, b1 - b0, c1 - c0
        g = g_dabi[:, a0:a1, b0:b1, :].reshape(v, na * nb_ * o)          # [d, (a b i)]
        t = t2[:, c0:c1, :, :].reshape(v, nc * o * o)                    # [d, (c j k)]
        x = g.T @ t                                                      # particle, K = v
        tl = t2_abil[a0:a1, b0:b1, :, :].reshape(na * nb_ * o, o)        # [(a b i), l]
        gl = g_aijk[c0:c1, :, :, :].reshape(nc * o * o, o)               # [(c j k), l]
        x -= tl @ gl.T                                                   # hole, K = o
        return x.reshape(na, nb_, o, nc, o, o)                           # a,b,i,c,j,k

    for ia, a in enumerate(vb):
        for ib, b in enumerate(vb[: ia + 1]):
            for ic, c in enumerate(vb[: ib + 1]):
                global_iter += 1
                if global_iter % size != rank:                # :477-480
                    continue
                if block_filter is not None and global_iter not in block_filter:
                    continue
                # six contraction pairs, each brought to (a,b,c,i,j,k) order (:486-557)
                t3 = np.einsum("abicjk->abcijk", t3_term(a, b, c))        # abcijk
                t3 = t3 + np.einsum("acibkj->abcijk", t3_term(a, c, b))   # acbikj
                t3 = t3 + np.einsum("cakbij->abcijk", t3_term(c, a, b))   # cabkij
                t3 = t3 + np.einsum("cbkaji->abcijk", t3_term(c, b, a))   # cbakji
                t3 = t3 + np.einsum("bcjaki->abcijk", t3_term(b, c, a))   # bcajki
                t3 = t3 + np.einsum("bajcik->abcijk", t3_term(b, a, c))   # bacjik
                # V (:559-596): v3(b,c,j,k,a,i) = g_abij(b,c,j,k) t1(a,i) and two more
                (a0, a1), (b0, b1), (c0, c1) = a, b, c
                v3 = np.einsum("bcjk,ai->abcijk", g_abij[b0:b1, c0:c1], t1[a0:a1])
                v3 = v3 + np.einsum("acik,bj->abcijk", g_abij[a0:a1, c0:c1], t1[b0:b1])
                v3 = v3 + np.einsum("abij,ck->abcijk", g_abij[a0:a1, b0:b1], t1[c0:c1])
                # :598-609
                z = (4.0 * t3 + _tr(t3, "abckij") + _tr(t3, "abcjki")
                     - 2.0 * (_tr(t3, "abckji") + _tr(t3, "abcikj") + _tr(t3, "abcjik")))
                result = (t3 + v3) * z
                offset = (a0, b0, c0, 0, 0, 0)
                if ib < ia and ic < ib:                                   # :619-627
                    e = 2.0 * reduce_plain(result, eps_occ, eps_vir, offset)
                else:                                                     # :629-638
                    e = reduce_symm(result, eps_occ, eps_vir, offset)
                e_total += e
                parts.append((global_iter, e))
    if return_parts:
        return e_total, parts
    return e_total


def coarse_block_count(v: int, vir_block: int = 8) -> int:
    nb = len(_blocks(v, vir_block))
    return nb * (nb + 1) * (nb + 2) // 6


# ----------------------------------------------------------------------------------------
# ijk-driven form (the north star's sharding; same sum re-ordered)
# ----------------------------------------------------------------------------------------
def triple_weight(i: int, j: int, k: int) -> float:
    """Weight of the ordered triple i>=j>=k: 2 all distinct, 1 two equal, 0 all equal.

    Same weights as CCSD_T_ReduceSymm (:2399-2423), applied to the occupied triple.
    """
    if i == j == k:
        return 0.0
    if i == j or j == k or i == k:
        return 1.0
    return 2.0


def ijk_triple_list(o: int):
    """Ordered list of (i,j,k), i>=j>=k, excluding i==j==k (weight 0).  This is the unit
    list the GPU scheduler shards; the enumeration order is part of the C-ABI contract
    (include/mpqc_t.h, mpqc_t_triple_count)."""
    return [(i, j, k) for i in range(o) for j in range(i + 1) for k in range(j + 1)
            if not (i == j == k)]


def w_ijk(t2, g_aijk, g_abci, i, j, k):
    """W^{abc}_{ijk} for one occupied triple as a [v,v,v] array: the six particle (K=v) and six
    hole (K=o) contractions of :314-340 / :1142-1146 evaluated at fixed (i,j,k)."""
    v = t2.shape[0]

    def x(i, j, k):
        # X[a,b,c] = sum_d g_dabi[d,a,b,i] t2[d,c,j,k] - sum_l g_cjkl[c,j,k,l] t2[a,b,i,l]
        g = g_abci[:, :, :, i].reshape(v, v * v)             # [d,(a b)]
        p = (g.T @ t2[:, :, j, k]).reshape(v, v, v)           # [(a b), c]
        h = (t2[:, :, i, :].reshape(v * v, -1) @ g_aijk[:, j, k, :].T).reshape(v, v, v)
        return p - h

    w = x(i, j, k)                                            # abc ; ijk
    w = w + x(i, k, j).transpose(0, 2, 1)                     # acb ; ikj
    w = w + x(k, i, j).transpose(1, 2, 0)                     # cab ; kij  -> X[c,a,b]
    w = w + x(k, j, i).transpose(2, 1, 0)                     # cba ; kji
    w = w + x(j, k, i).transpose(2, 0, 1)                     # bca ; jki  -> X[b,c,a]
    w = w + x(j, i, k).transpose(1, 0, 2)                     # bac ; jik
    return w


def v_ijk(t1, g_abij, i, j, k):
    """V^{abc}_{ijk}, :1150-1152 at fixed (i,j,k)."""
    return (np.einsum("ab,c->abc", g_abij[:, :, i, j], t1[:, k])
            + np.einsum("bc,a->abc", g_abij[:, :, j, k], t1[:, i])
            + np.einsum("ac,b->abc", g_abij[:, :, i, k], t1[:, j]))


def energy_ijk(t1, t2, g_abij, g_aijk, g_abci, eps_occ, eps_vir, i, j, k) -> float:
    """sum_abc (W+V) Z / D at fixed (i,j,k) (unweighted)."""
    w = w_ijk(t2, g_aijk, g_abci, i, j, k)
    vv = v_ijk(t1, g_abij, i, j, k)
    # Z[a,b,c] = 4W[abc] + W[bca] + W[cab] - 2(W[cba] + W[acb] + W[bac])   
    z = (4.0 * w + w.transpose(2, 0, 1) + w.transpose(1, 2, 0)
         - 2.0 * (w.transpose(2, 1, 0) + w.transpose(0, 2, 1) + w.transpose(1, 0, 2)))
    ev = eps_vir
    d = (eps_occ[i] + eps_occ[j] + eps_occ[k]
         - ev[:, None, None] - ev[None, :, None] - ev[None, None, :])
    return float(np.sum((w + vv) * z / d))


def ijk_driven(t1, t2, g_abij, g_aijk, g_abci, eps_occ, eps_vir, triples=None,
               return_parts: bool = False):
    """E(T) = sum_{i>=j>=k} w_ijk sum_abc (W+V) Z / D."""
    o = len(eps_occ)
    if triples is None:
        triples = ijk_triple_list(o)
    parts = []
    for (i, j, k) in triples:
        e = triple_weight(i, j, k) * energy_ijk(t1, t2, g_abij, g_aijk, g_abci,
                                                eps_occ, eps_vir, i, j, k)
        parts.append(e)
    e_total = float(np.sum(np.asarray(parts))) if parts else 0.0
    if return_parts:
        return e_total, np.asarray(parts)
    return e_total


def vblock_energies(t1, t2, g_abij, g_aijk, g_abci, eps_occ, eps_vir, vir_block: int = 8):
    """The SAME energy decomposed over virtual-block triples a >= b >= c but accumulated in the ijk-driven order -- the
    identity behind ``mpqc_t_run_vblocks`` (include/mpqc_t.h):

        e_block[tt] = sum_{i>=j>=k} w_ijk * sum_{(a,b,c) in orbit(tt)} (W+V) Z / D

    where orbit(tt) is the union of the DISTINCT permutations of the three tiles (TA, TB, TC).  Because
    sum_{ijk} f(abc, ijk) is symmetric in (a,b,c) and the orbit is closed under permutations, e_block[tt] equals what
    iteration ``global_iter = tt + 1`` of the reference's coarse loop adds (ccsd_t.h:443-480, :619-638), including the
    CCSD_T_ReduceSymm weights of the blocks with coinciding tiles.  Returns the list in global_iter order."""
    o, v = len(eps_occ), len(eps_vir)
    vb = _blocks(v, vir_block)
    sets = [(a, b, c) for a in range(len(vb)) for b in range(a + 1) for c in range(b + 1)]
    out = np.zeros(len(sets))
    for (i, j, k) in ijk_triple_list(o):
        w = w_ijk(t2, g_aijk, g_abci, i, j, k)
        vv = v_ijk(t1, g_abij, i, j, k)
        z = (4.0 * w + w.transpose(2, 0, 1) + w.transpose(1, 2, 0)
             - 2.0 * (w.transpose(2, 1, 0) + w.transpose(0, 2, 1) + w.transpose(1, 0, 2)))
        d = (eps_occ[i] + eps_occ[j] + eps_occ[k]
             - eps_vir[:, None, None] - eps_vir[None, :, None] - eps_vir[None, None, :])
        f = (w + vv) * z / d
        wt = triple_weight(i, j, k)
        for tt, tiles in enumerate(sets):
            acc = 0.0
            for perm in set(itertools.permutations(tiles)):
                (a0, a1), (b0, b1), (c0, c1) = vb[perm[0]], vb[perm[1]], vb[perm[2]]
                acc += float(f[a0:a1, b0:b1, c0:c1].sum())
            out[tt] += wt * acc
    return out
