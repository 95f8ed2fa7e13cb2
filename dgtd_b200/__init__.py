"""dgtd_b200 — B200-native DG-Maxwell evolution hot path (OpenSEMBA/dgtd drop-in).

The product is `libdgtd_b200.so` (hand-written sm_100a CUDA kernels behind the C ABI of
include/dgtd_b200.h) plus the header-only C++ shells in dgtd_b200/mfem_shell/B200Evolution.h.  This Python package is a thin
ctypes binding used by tests/ and bench.py; it contains no numerics and NO CPU fallback — loading
fails loudly if the shared library has not been built (see __graft_entry__.build()).
"""
from .api import (BC_NONE, BC_PEC, BC_PMC, BC_SMA, DgtdError, Evolution, Gather, Mesh, PlaneWave, lib, lib_path,
                  setup_query, HEADER_SYMBOLS)

__all__ = ["BC_NONE", "BC_PEC", "BC_PMC", "BC_SMA", "DgtdError", "Evolution", "Gather", "Mesh", "PlaneWave", "lib",
           "lib_path", "setup_query", "HEADER_SYMBOLS"]
