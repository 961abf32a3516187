"""mpqc_b200: B200-native perturbative-triples (T) energy for MPQC4's CCSD(T).

Only what the hot path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the C ABI of
``include/mpqc_t.h``), ``lib`` (ctypes marshalling), ``ccsd_t`` (host-side mirror of the reference's
``CCSD_T`` interface) and ``synthetic`` (symmetry-correct synthetic inputs).
"""
__version__ = "0.1.0"
