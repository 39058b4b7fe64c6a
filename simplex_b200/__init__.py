"""simplex_b200 — B200-native implementation of the per-iteration hot path of
maciejdrwal/simplex's revised simplex (reference src/Simplex.cpp:255-353, Basis.cpp,
InitialBasis.cpp) behind a C ABI (include/simplex_b200.h).

    csrc/     hand-written sm_100a CUDA kernels + the C ABI  -> libsimplex_b200.so
    host/     C++ host side mirroring the reference's classes (LinearProgram, Parser, Presolve,
              InitialBasis, Simplex, Basis) and the `simplex <file.lp>` CLI
    solver.py ctypes mirror of the seam for Python callers, tests and bench.py
    synth.py  synthetic dense LP generator G(m,n,n_eq,seed) used by the benchmark configs
"""
from .solver import Basis, DeviceSimplex, algorithmic_bytes_per_pivot  # noqa: F401
from ._abi import Sb200Error  # noqa: F401
