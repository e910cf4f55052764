"""B200-native ensemble ODE integrator with the MPGOS user-facing surface.

The product is the CUDA library ``libmpgos_b200.so`` (C-ABI in ``include/mpgos_b200.h``) and the
C++ drop-in header ``include/SingleSystem_PerThread_Interface.cuh``.  This package is the thin
Python binding used by the tests and the bench: :class:`ProblemSolver` mirrors the reference's
host object method for method on top of the C-ABI.  There is no CPU fallback -- importing works
anywhere (so that CPU-only checks can inspect the library's symbols), but creating a solver
without a CUDA device raises.
"""
from .solver import (ProblemSolver, MpgosError, load_library, library_path, build_library,
                     Variables, Options, Algorithms, run_workload, ABI_SYMBOLS, probe_fma_peak, probe_dmma_peak, fast_math_probe,
                     apply_workload, read_back)
from . import workloads
from . import workloads_coupled
from .solver_coupled import (CoupledProblemSolver, CoupledVariables, CoupledOptions, COUPLED_ABI_SYMBOLS, run_coupled_workload,
                             apply_coupled_workload, read_back_coupled)

__all__ = ["ProblemSolver", "MpgosError", "load_library", "library_path", "build_library", "Variables",
           "Options", "Algorithms", "run_workload", "workloads", "ABI_SYMBOLS", "probe_fma_peak", "probe_dmma_peak", "apply_workload", "read_back"]
