"""Host-side mirror of the reference's ``CCSD_T`` wavefunction for the (T) path.

Mirrors ``mpqc::lcao::CCSD_T<Tile,Policy>`` (/root/reference/src/mpqc/chemistry/qc/lcao/cc/ccsd_t.h:33-197):
same KeyVal keywords, same method names (``compute_ccsd_t``, ``evaluate``, ``triples_energy``,
``obsolete``), same error behaviour (``InputError`` on an invalid ``approach``), same log lines --
but the body of ``compute_ccsd_t`` hands the dense blocks through the C ABI of
``include/mpqc_t.h`` to the sm_100a kernels.  The compiled, MPQC-facing adapter with identical
structure is the in-class patch ``integration/mpqc_ccsd_t_gpu.patch`` + ``integration/ccsd_t_gpu_impl.h``; this Python mirror exists so that tests and
``bench.py`` exercise the same boundary without TiledArray/MADWorld.

The CCSD base class of the reference (``ccsd.h:54-265``) is *not* rebuilt: it is represented by a
"provider" object that supplies what ``CCSD_T`` reads from it:

    t1() [v,o], t2() [v,v,o,o]                      ccsd.h:165-179
    orbital_energy()  (all MOs, frozen core first)  ccsd.h:141-148
    trange1_engine() -> get_occ(), get_nfrozen(), get_active_occ(), get_vir()   trange1_engine.h:60-72
    get_abij(), get_aijk(), get_abci()              ccsd_t.h:2210-2244 (post-permutation layouts)
    ccsd_energy()                                   value CCSD::evaluate leaves in the Energy result
"""
from __future__ import annotations

import ctypes as C
import sys
import time
from dataclasses import dataclass, field

import numpy as np

from . import lib as L


class InputError(ValueError):
    """mpqc::InputError (util/core/exception.h): bad user input; carries the offending keyword."""

    def __init__(self, msg, keyword=None, value=None):
        super().__init__(msg)
        self.keyword = keyword
        self.value = value


class FeatureDisabled(RuntimeError):
    """mpqc::FeatureDisabled -> exit code 2 in mpqc.cpp:261-264: raised when no CUDA device exists."""


class MemAllocFailed(MemoryError):
    """mpqc::MemAllocFailed"""


class ProgrammingError(RuntimeError):
    """mpqc::ProgrammingError"""


_STATUS_TO_EXC = {
    L.ERR_BAD_ARG: InputError, L.ERR_NO_DEVICE: FeatureDisabled, L.ERR_OOM: MemAllocFailed,
    L.ERR_CUDA: ProgrammingError, L.ERR_NCCL: ProgrammingError, L.ERR_INTERNAL: ProgrammingError,
}


def _raise_for(status: int, what: str):
    if status == L.OK:
        return
    lib = L.load()
    detail = lib.mpqc_t_strerror(status).decode() + "; " + lib.mpqc_t_last_error().decode()
    raise _STATUS_TO_EXC.get(status, ProgrammingError)(f"{what}: {detail}")


class TRange1Engine:
    """Counts only (expression/trange1_engine.h:60-72); the GPU path needs no tiling."""

    def __init__(self, n_occ: int, n_all: int, n_frozen: int = 0):
        self._occ, self._all, self._nfrozen = int(n_occ), int(n_all), int(n_frozen)

    def get_occ(self):
        return self._occ

    def get_nfrozen(self):
        return self._nfrozen

    def get_active_occ(self):
        return self._occ - self._nfrozen

    def get_vir(self):
        return self._all - self._occ


@dataclass
class DenseCCSD:
    """A converged-CCSD provider holding dense arrays (numpy on the host, or torch CUDA tensors).

    ``orbital_energies`` covers all MOs (frozen core, active occupied, virtual) exactly like the
    reference's ``orbital_energy()``; the (T) driver slices it.
    """
    t1_: object
    t2_: object
    g_abij: object
    g_aijk: object
    g_abci: object
    orbital_energies: object
    n_frozen: int = 0
    e_ccsd: float = 0.0
    # optional density-fitting factors (CCSD::get_Xab / get_Xij / get_Xai, ccsd.h:480-493); when present and
    # is_df() the (T) driver hands THEM to the library and the v^3 o tensor is assembled on the device
    x_ab: object = None
    x_ij: object = None
    x_ai: object = None
    _engine: TRange1Engine = field(default=None, repr=False)

    def __post_init__(self):
        v, o = self.t1_.shape
        n_all = len(self.orbital_energies)
        n_occ = self.n_frozen + o
        if n_all != n_occ + v:
            raise InputError("orbital_energies must cover frozen + active occupied + virtual orbitals")
        self._engine = TRange1Engine(n_occ, n_all, self.n_frozen)

    def t1(self):
        return self.t1_

    def t2(self):
        return self.t2_

    def orbital_energy(self):
        return self.orbital_energies

    def trange1_engine(self):
        return self._engine

    def get_abij(self):
        return self.g_abij

    def get_aijk(self):
        return self.g_aijk

    def get_abci(self):
        return self.g_abci

    def ccsd_energy(self):
        return self.e_ccsd

    def is_df(self):                                   # ccsd.h:93-99: method == "df"
        return self.x_ab is not None and self.x_ij is not None and self.x_ai is not None

    def get_Xab(self):
        return self.x_ab

    def get_Xij(self):
        return self.x_ij

    def get_Xai(self):
        return self.x_ai

    @classmethod
    def from_problem(cls, p: dict, n_frozen: int = 0, e_ccsd: float = 0.0, frozen_eps=None):
        """Wrap a ``synthetic.make_problem`` dict (optionally prepending frozen-core energies)."""
        is_np = isinstance(p["eps_occ"], np.ndarray)
        if n_frozen:
            fe = frozen_eps if frozen_eps is not None else np.linspace(-20.0, -10.0, n_frozen)
            if is_np:
                eps = np.concatenate([np.asarray(fe, dtype=np.float64), p["eps_occ"], p["eps_vir"]])
            else:
                import torch
                eps = torch.cat([torch.as_tensor(fe, dtype=torch.float64, device=p["eps_occ"].device),
                                 p["eps_occ"], p["eps_vir"]])
        else:
            if is_np:
                eps = np.concatenate([p["eps_occ"], p["eps_vir"]])
            else:
                import torch
                eps = torch.cat([p["eps_occ"], p["eps_vir"]])
        return cls(p["t1"], p["t2"], p["g_abij"], p["g_aijk"], p["g_abci"], eps, n_frozen, e_ccsd,
                   p.get("x_ab"), p.get("x_ij"), p.get("x_ai"))


class Energy:
    """Minimal stand-in for the Energy property's result slot (properties/energy.h:14-34)."""

    def __init__(self):
        self.value = None


_VALID_APPROACH = ("coarse", "fine", "straight", "laplace", "gpu")


class CCSD_T:
    """Drop-in mirror of ``CCSD_T`` behind the factory key ``"type": "CCSD(T)"`` (ccsd_t.cpp:10,13).

    Keywords (ccsd_t.h:89-97,102-131): ``approach`` (coarse|fine|straight|laplace, plus the new
    default ``gpu``), ``increase``, ``reblock_occ``, ``reblock_unocc``, ``reblock_inner``,
    ``replicate_ijka``, ``quadrature_points``; all accepted so existing inputs keep working.  The CPU
    blocking keywords are parsed and ignored: coarse/fine/straight are the same exact sum and all run
    on the GPU path.  ``laplace`` is a different (approximate) method, out of scope here, and raises
    ``FeatureDisabled``.

    Extra keywords of the GPU path: ``ngpu`` (devices driven by this process), ``device_ids``,
    ``batch``, ``use_nccl`` (one-shot NCCL communicator for ``ngpu`` > 1).  Multi-rank use, one rank per GPU
    (replaces ``gop.sum`` at ccsd_t.h:692), either
      * ``comm=``: a library communicator (``mpqc_t_comm_create_rank`` / ``_local``, a ``ctypes.c_void_p``); the call
        becomes ``mpqc_t_energy_comm``: inputs replicated over NVLink, unit energies summed by ncclAllReduce inside
        the library, every rank gets the TOTAL E(T); or
      * ``rank`` / ``world_size`` keywords + ``reduce=``: the unit list is sharded by stride and ``reduce(partial)``
        (the host program's own sum, e.g. an MPI/torch all-reduce) must return the total.  A sharded run without
        either raises ``InputError`` -- a partial E(T) is never stored as the energy.
    """

    def __init__(self, kv: dict, ccsd=None, out=None, comm=None, reduce=None):
        kv = dict(kv or {})
        t = kv.get("type", "CCSD(T)")
        if t != "CCSD(T)":
            raise InputError(f"CCSD_T constructed from a KeyVal of type {t!r}", "type", t)
        self._ccsd = ccsd if ccsd is not None else kv.get("ccsd")
        self.reblock_ = ("reblock_occ" in kv) or ("reblock_unocc" in kv)          # :102
        self.occ_block_size_ = int(kv.get("reblock_occ", 8))                      # :104
        self.unocc_block_size_ = int(kv.get("reblock_unocc", 8))                  # :105
        self.replicate_ijka_ = bool(kv.get("replicate_ijka", False))              # :106
        self.inner_block_size_ = int(kv.get("reblock_inner", 0)) if self.reblock_ else 0
        self.reblock_inner_ = self.inner_block_size_ != 0
        self.increase_ = int(kv.get("increase", 2))                               # :116
        self.approach_ = str(kv.get("approach", "gpu"))
        if self.approach_ not in _VALID_APPROACH:                                 # :118-122
            raise InputError("Invalid (T) approach! \n", "approach", self.approach_)
        if self.approach_ == "laplace":                                           # :125-128
            self.reblock_ = False
            self.reblock_inner_ = False
        self.n_laplace_quad_ = int(kv.get("quadrature_points", 4))                # :131
        self.verbose_ = bool(kv.get("verbose", False))
        # GPU-path keywords
        # (the in-class patch of the reference spells them ngpu / gpu_batch / gpu_df / gpu_dump_file; both spellings work)
        self.ngpu_ = int(kv.get("ngpu", 1))
        self.device_ids_ = kv.get("device_ids")
        self.batch_ = int(kv.get("gpu_batch", kv.get("batch", 0)))
        self.use_nccl_ = bool(kv.get("use_nccl", False))
        # "gpu_df" / "df_direct": with a density-fitted CCSD (method df) assemble the integrals on the device from the
        # three-centre factors instead of receiving the dense <ia|bc> tensor.  Off by default in this
        # mirror so that tests choose the input form explicitly (the patched reference defaults to on when is_df()).
        self.df_direct_ = bool(kv.get("gpu_df", kv.get("df_direct", False)))
        self.df_block_ = int(kv.get("gpu_df_block", 0))      # 0 automatic, -1 resident, b > 0 panel cache (mpqc_t_options.df_block)
        self.dump_file_ = str(kv.get("gpu_dump_file", ""))
        self.rank_ = int(kv.get("rank", 0))
        self.world_size_ = int(kv.get("world_size", 1))
        if self.ngpu_ < 1:
            raise InputError("ngpu must be >= 1", "ngpu", self.ngpu_)
        if not (0 <= self.rank_ < self.world_size_):
            raise InputError("rank must satisfy 0 <= rank < world_size", "rank", self.rank_)
        self._comm = comm
        self._reduce = reduce
        if self._comm is not None and self.world_size_ != 1:
            raise InputError("give either a library communicator (comm=) or rank/world_size + reduce=, not both", "rank",
                             self.rank_)
        self.triples_energy_ = 0.0
        self.computed_ = False
        self.stats_ = None
        self._out = out if out is not None else sys.stdout

    # -- Wavefunction protocol -------------------------------------------------------------
    def obsolete(self):                                                           # :136-139
        self.triples_energy_ = 0.0
        self.computed_ = False

    def triples_energy(self) -> float:                                            # :141
        return self.triples_energy_

    def computed(self) -> bool:
        return self.computed_

    def can_evaluate(self, energy) -> bool:                                       # ccsd.h:196-199
        return isinstance(energy, Energy)

    def stats(self):
        return self.stats_

    def evaluate(self, result: Energy):                                           # :179-197
        if not self.computed_:
            ccsd_energy = float(self._ccsd.ccsd_energy())
            t0 = time.perf_counter()
            self.compute_ccsd_t()
            self.computed_ = True
            result.value = ccsd_energy + self.triples_energy_
            print(f"(T) Time in CCSD(T): {time.perf_counter() - t0} S", file=self._out)
        return result

    # -- the hot path ------------------------------------------------------------------------
    def compute_ccsd_t(self) -> float:                                            # :144-177
        if self._ccsd is None:
            raise ProgrammingError("CCSD_T has no converged CCSD provider")
        if self.approach_ == "laplace":
            raise FeatureDisabled("approach=laplace is a different (approximate) method; "
                                  "not provided by the GPU (T) path")
        if self.world_size_ > 1 and self._reduce is None:
            raise InputError("rank/world_size sharding needs reduce= (or use a library communicator, comm=): the partial "
                             "E(T) of one rank must not be stored as the energy", "world_size", self.world_size_)
        t0 = time.perf_counter()
        print("\nBegining CCSD(T) ", file=self._out)
        cc = self._ccsd
        eng = cc.trange1_engine()
        n_occ, n_frozen = eng.get_occ(), eng.get_nfrozen()
        o, v = eng.get_active_occ(), eng.get_vir()
        eps = cc.orbital_energy()
        eps_occ = eps[n_frozen:n_occ]                # ccsd_t.h:2306-2311: eps[i + n_frozen]
        eps_vir = eps[n_occ:n_occ + v]               # eps[a + n_occ]
        on_device = not isinstance(eps, np.ndarray) and getattr(eps, "is_cuda", False)
        if on_device:
            import torch
            torch.cuda.current_stream().synchronize()    # producers of the device tensors (the library syncs the device too)
        if isinstance(eps, np.ndarray):
            eps_occ = np.ascontiguousarray(eps_occ, dtype=np.float64)
            eps_vir = np.ascontiguousarray(eps_vir, dtype=np.float64)
        else:
            eps_occ, eps_vir = eps_occ.contiguous(), eps_vir.contiguous()
        use_df = self.df_direct_ and hasattr(cc, "is_df") and cc.is_df()
        if self.df_direct_ and not use_df:
            raise InputError("df_direct requested but the CCSD provider has no density-fitting factors", "df_direct")
        if use_df:
            naux = cc.get_Xab().shape[0]
            prob = L.make_df_problem(o, v, naux, eps_occ, eps_vir, cc.t1(), cc.t2(), cc.get_Xab(), cc.get_Xij(),
                                     cc.get_Xai())
        else:
            prob = L.make_problem(o, v, eps_occ, eps_vir, cc.t1(), cc.t2(), cc.get_abij(), cc.get_aijk(),
                                  cc.get_abci())
        opt = L.Options()
        opt.ngpu = self.ngpu_
        if self.device_ids_ is not None:
            ids = (C.c_int32 * len(self.device_ids_))(*self.device_ids_)
            opt.device_ids = ids
        opt.verbose = 0
        opt.inputs_on_device = 1 if on_device else 0
        opt.unit_first = self.rank_
        opt.unit_stride = self.world_size_
        opt.unit_count = -1
        opt.batch = self.batch_
        opt.df_block = self.df_block_
        opt.use_nccl = 1 if self.use_nccl_ else 0
        if self.dump_file_ and not use_df and isinstance(eps, np.ndarray):
            from . import dump                                  # same MPQCT001 file the patched reference writes
            dump.save_problem(self.dump_file_, eps, n_frozen, cc.t1(), cc.t2(), cc.get_abij(), cc.get_aijk(), cc.get_abci())
        st = L.Stats()
        e = C.c_double(0.0)
        lib = L.load()
        if self._comm is not None:
            opt.unit_first, opt.unit_stride = 0, 1             # the communicator shards the whole unit list
            fn = lib.mpqc_t_energy_df_comm if use_df else lib.mpqc_t_energy_comm
            status = fn(self._comm, C.byref(prob), C.byref(opt), C.byref(e), C.byref(st))
        elif use_df:
            status = lib.mpqc_t_energy_df(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st))
        else:
            status = lib.mpqc_t_energy(C.byref(prob), C.byref(opt), C.byref(e), C.byref(st))
        _raise_for(status, "mpqc_t_energy_df" if use_df else "mpqc_t_energy")
        # the partial of a stride-sharded run goes through the host program's sum before it becomes the energy
        self.triples_energy_ = float(self._reduce(e.value)) if self.world_size_ > 1 else e.value
        self.stats_ = st.as_dict()
        print(f"(T) Energy: {self.triples_energy_} Time: {time.perf_counter() - t0} S ", file=self._out)
        return self.triples_energy_


class CCSD_T_F12(CCSD_T):
    """Mirror of the path's SECOND caller, ``CCSD_T_F12<Tile>::evaluate`` (f12/ccsd_t_f12.h:47-69): it computes the
    CCSD(F12) energy first (out of scope here: supplied by the provider), purges the integral registries, and then calls
    the protected, non-virtual ``CCSD_T::compute_ccsd_t()`` of its base and adds ``triples_energy()``.  Because the GPU
    path is dispatched INSIDE ``compute_ccsd_t`` (integration/mpqc_ccsd_t_gpu.patch), this caller gets it unchanged."""

    def __init__(self, kv: dict, ccsd=None, out=None, comm=None, reduce=None):
        kv = dict(kv or {})
        t = kv.pop("type", "CCSD(T)F12")
        if t != "CCSD(T)F12":
            raise InputError(f"CCSD_T_F12 constructed from a KeyVal of type {t!r}", "type", t)
        super().__init__(dict(kv, type="CCSD(T)"), ccsd=ccsd, out=out, comm=comm, reduce=reduce)

    def evaluate(self, result: Energy):                                           # f12/ccsd_t_f12.h:47-69
        if not self.computed_:
            cc = self._ccsd
            ccsd_f12_energy = float(cc.ccsd_f12_energy() if hasattr(cc, "ccsd_f12_energy") else cc.ccsd_energy())
            t0 = time.perf_counter()
            self.compute_ccsd_t()                                                 # :59
            print(f"(T) Time in CCSD(T)F12:  {time.perf_counter() - t0}", file=self._out)
            self.computed_ = True
            result.value = ccsd_f12_energy + self.triples_energy()                # :67
        return result


#: the KeyVal registry entries of the drop-in (keyval.h:129-139 asserts one registration per key)
REGISTRY = {"CCSD(T)": CCSD_T, "CCSD(T)F12": CCSD_T_F12}


def class_ptr(kv: dict, **deps):
    """KeyVal::class_ptr analogue (keyval.h:865-944): construct the object named by ``kv['type']``."""
    t = kv.get("type")
    if t not in REGISTRY:
        raise InputError(f"KeyVal type {t!r} is not registered", "type", t)
    return REGISTRY[t](kv, **deps)
