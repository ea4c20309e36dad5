"""blasted_b200: B200-native asynchronous ILU(0)/SGS preconditioners, level-scheduled variants and
CSR/BSR SpMV behind the interface of Slaedr/BLASTed's SRPreconditioner / SRFactory objects.

The compute path is hand-written CUDA for sm_100a in blasted_b200/csrc, exposed through the C ABI of
include/blasted_b200.h (libblasted_b200.so).  This package is the thin Python mirror of the
reference's C++ host interface (include/solverfactory.hpp, solverops_base.hpp, blockmatrices.hpp,
tests/solvers.hpp); the C++ mirror is blasted_b200/host/b200_solverops.hpp.
"""
from . import matgen                                         # noqa: F401
from .solverfactory import (SRFactory, AsyncSolverSettings, Preconditioner, SRMatrixView,   # noqa: F401
                            CSRMatrixView, BSRMatrixView, PrecInfo, SolveInfo, BiCGSTAB, GCR, FGMRES,
                            RichardsonSolver, device_count, kernel_launches,
                            reset_kernel_launches, SOLVER_TYPES)
