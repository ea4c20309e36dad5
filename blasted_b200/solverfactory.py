"""Python mirror of the reference's host interface for the asynchronous-preconditioner path.

Same names, argument meaning and error behaviour as the C++ originals:

  SRFactory.create_preconditioner / solverTypeFromString   include/solverfactory.hpp:71-124,
                                                            src/solverfactory.cpp:34-228
  AsyncSolverSettings                                       include/solverfactory.hpp:46-68
  Preconditioner.dim/compute/apply/apply_relax/
      relaxationAvailable/setApplyParams                    include/solverops_base.hpp:32-64
  CSRMatrixView / BSRMatrixView .apply/.gemv3/.dim          include/blockmatrices.hpp:71-160
  BiCGSTAB / GCR / RichardsonSolver .setParams/.solve       tests/solvers.hpp:29-135
  PrecInfo                                                  include/preconditioner_diagnostics.hpp:14-41

Every call goes through the C ABI (libblasted_b200.so); nothing is computed in Python.  numpy
arrays are treated as the reference treats raw host pointers (copied to/from the device inside the
call); torch CUDA tensors are passed as device pointers and stay on the device.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import lib, check, Settings
from .matgen import SRMatrix

# include/solvertypes.h:14-26 and the strings of include/solverfactory.hpp:22-43
SOLVER_TYPES = {"jacobi": 0, "gs": 1, "sgs": 2, "ilu0": 3, "seqilu0": 4, "sfilu0": 5, "sapilu0": 6,
                "cscbgs": 7, "level_sgs": 8, "async_level_ilu0": 9, "none": 10}
FACT_INIT = {"init_zero": 0, "init_original": 1, "init_sgs": 2, "init_none": 3}
APPLY_INIT = {"init_zero": 0, "init_jacobi": 1, "init_none": 2}
COLMAJOR, ROWMAJOR = 0, 1
LEVELS_DAG, LEVELS_CONTIGUOUS = 0, 1


def getFactInitFromString(s: str) -> int:
    """include/async_initialization_decl.hpp:38-49"""
    if s not in FACT_INIT:
        raise ValueError("Factor initialization not recongnized!")
    return FACT_INIT[s]


def getApplyInitFromString(s: str) -> int:
    """include/async_initialization_decl.hpp:52-61"""
    if s not in APPLY_INIT:
        raise ValueError("Apply initialization not recongnized!")
    return APPLY_INIT[s]


def device_count() -> int:
    return lib.b200_device_count()


def kernel_launches() -> int:
    return lib.b200_kernel_launches()


def reset_kernel_launches() -> None:
    lib.b200_reset_kernel_launches()


@dataclass
class AsyncSolverSettings:
    """include/solverfactory.hpp:46-68 (SolverSettings + AsyncSolverSettings)."""
    prectype: int = SOLVER_TYPES["ilu0"]
    bs: int = 1
    blockstorage: int = COLMAJOR
    relax: bool = False
    thread_chunk_size: int = 0
    scale: bool = False
    nbuildsweeps: int = 1
    napplysweeps: int = 1
    fact_inittype: int = FACT_INIT["init_original"]
    apply_inittype: int = APPLY_INIT["init_jacobi"]
    compute_precinfo: bool = False
    level_mode: int = LEVELS_DAG          # device-only knob, see include/blasted_b200.h

    def to_c(self) -> Settings:
        return Settings(int(self.prectype), int(self.bs), int(self.blockstorage), int(self.relax),
                        int(self.thread_chunk_size), int(self.scale), int(self.nbuildsweeps),
                        int(self.napplysweeps), int(self.fact_inittype), int(self.apply_inittype),
                        int(self.compute_precinfo), int(self.level_mode))


@dataclass
class PrecInfo:
    """include/preconditioner_diagnostics.hpp:14-41"""
    f_info: np.ndarray = field(default_factory=lambda: np.zeros(6))

    def prec_remainder_norm(self): return self.f_info[0]
    def prec_rem_initial_norm(self): return self.f_info[1]
    def upper_min_diag_dom(self): return self.f_info[2]
    def upper_avg_diag_dom(self): return self.f_info[3]
    def lower_min_diag_dom(self): return self.f_info[4]
    def lower_avg_diag_dom(self): return self.f_info[5]


@dataclass
class SolveInfo:
    """tests/solvers.hpp:19-27"""
    converged: bool = False
    iters: int = 0
    resnorm: float = 0.0
    bnorm: float = 0.0
    walltime: float = 0.0          # seconds of device time (CUDA events)
    precapplywtime: float = 0.0


def _is_torch_cuda(a) -> bool:
    return type(a).__module__.startswith("torch") and getattr(a, "is_cuda", False)


def _host(a, writable=False):
    a = np.ascontiguousarray(a, dtype=np.float64) if not writable else a
    if writable and (a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]):
        raise ValueError("output array must be contiguous float64")
    return a, a.ctypes.data_as(C.c_void_p)


class SRMatrixView:
    """Device-resident operator: CSRMatrixView / BSRMatrixView (include/blockmatrices.hpp:71-160).

    Unlike the reference's views (which wrap host arrays), construction copies the matrix to HBM.
    """

    def __init__(self, m: SRMatrix):
        self.m = m
        self._h = C.c_void_p()
        di = m.diagind.ctypes.data_as(C.c_void_p) if m.diagind is not None else None
        check(lib.b200_mat_create_host(m.nbrows, m.bs, ROWMAJOR if m.rowmajor else COLMAJOR,
                                       m.browptr.ctypes.data_as(C.c_void_p),
                                       m.bcolind.ctypes.data_as(C.c_void_p),
                                       m.vals.ctypes.data_as(C.c_void_p), di, C.byref(self._h)))

    @classmethod
    def from_device(cls, nbrows: int, bs: int, browptr, bcolind, vals, rowmajor: bool = False,
                    keep_host_copy: bool = False) -> "SRMatrixView":
        """Matrix assembled on the device (torch CUDA tensors: int32 browptr / bcolind, float64
        vals in the caller's block layout); the arrays are copied, diagonals are located on the
        device.  No host copy is kept unless asked for (`self.m` is then None)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self._bs, self._rowmajor = bs, rowmajor
        self.m = None
        import torch
        assert browptr.dtype == torch.int32 and bcolind.dtype == torch.int32 and vals.dtype == torch.float64
        check(lib.b200_mat_create_device(nbrows, bs, ROWMAJOR if rowmajor else COLMAJOR,
                                         C.c_void_p(browptr.data_ptr()), C.c_void_p(bcolind.data_ptr()),
                                         C.c_void_p(vals.data_ptr()), C.byref(self._h)))
        if keep_host_copy:
            self.m = self.to_host()
        return self

    @classmethod
    def from_handle(cls, handle, bs: int, rowmajor: bool) -> "SRMatrixView":
        """Wraps a matrix that was built on the device (front end); the host copy is fetched."""
        self = cls.__new__(cls)
        self._h = handle
        self._bs, self._rowmajor = bs, rowmajor
        self.m = None
        self.m = self.to_host()
        return self

    def to_host(self) -> SRMatrix:
        """The resident matrix as host arrays in the reference's raw layout."""
        bs = self.m.bs if self.m is not None else self._bs
        rowmajor = self.m.rowmajor if self.m is not None else self._rowmajor
        nb, nnzb = lib.b200_mat_nbrows(self._h), lib.b200_mat_nnzb(self._h)
        browptr = np.zeros(nb + 1, dtype=np.int32)
        bcolind = np.zeros(nnzb, dtype=np.int32)
        diagind = np.zeros(nb, dtype=np.int32)
        vals = np.zeros(nnzb * bs * bs, dtype=np.float64)
        check(lib.b200_mat_get_host(self._h, browptr.ctypes.data_as(C.c_void_p),
                                    bcolind.ctypes.data_as(C.c_void_p),
                                    diagind.ctypes.data_as(C.c_void_p),
                                    vals.ctypes.data_as(C.c_void_p)))
        return SRMatrix(nb, bs, browptr, bcolind, vals, diagind, rowmajor)

    def dim(self) -> int:
        return lib.b200_mat_dim(self._h)

    def update_values(self, vals) -> None:
        if _is_torch_cuda(vals):
            check(lib.b200_mat_update_values_device(self._h, C.c_void_p(vals.data_ptr())))
        else:
            _, p = _host(vals)
            check(lib.b200_mat_update_values_host(self._h, p))

    def apply(self, x, y=None):
        """y = A x  (AbstractLinearOperator::apply, include/linearoperator.hpp:36)"""
        if _is_torch_cuda(x):
            import torch
            y = torch.empty_like(x) if y is None else y
            check(lib.b200_mat_apply(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
            return y
        xa, xp = _host(x)
        y = np.empty(self.dim()) if y is None else y
        _, yp = _host(y, True)
        check(lib.b200_mat_apply_host(self._h, xp, yp))
        return y

    def gemv3(self, a, x, b, y, z=None):
        """z = a A x + b y  (MatrixView::gemv3, include/linearoperator.hpp:125-131)"""
        if _is_torch_cuda(x):
            import torch
            z = torch.empty_like(x) if z is None else z
            check(lib.b200_mat_gemv3(self._h, a, C.c_void_p(x.data_ptr()), b,
                                     C.c_void_p(y.data_ptr()), C.c_void_p(z.data_ptr())))
            return z
        _, xp = _host(x)
        ya, yp = _host(y)
        z = np.empty(self.dim()) if z is None else z
        _, zp = _host(z, True)
        check(lib.b200_mat_gemv3_host(self._h, a, xp, b, yp, zp))
        return z

    def set_stream(self, stream_ptr: int) -> None:
        lib.b200_mat_set_stream(self._h, C.c_void_p(stream_ptr))

    def release_workspace(self) -> None:
        """Frees the Krylov basis storage the solvers keep with the operator between solves."""
        check(lib.b200_mat_release_workspace(self._h))

    def close(self):
        if self._h:
            lib.b200_mat_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CSRMatrixView(SRMatrixView):
    def __init__(self, m: SRMatrix):
        if m.bs != 1:
            raise ValueError("CSRMatrixView needs a scalar matrix")
        super().__init__(m)


class BSRMatrixView(SRMatrixView):
    pass


class Preconditioner:
    """SRPreconditioner on the device (include/solverops_base.hpp:32-78).

    Created by SRFactory.create_preconditioner; `compute()` builds it from the current matrix
    values, `apply(r, z)` computes z = M^-1 r, `apply_relax(b, x)` relaxes A x = b in place.
    """

    def __init__(self, view: SRMatrixView, settings: AsyncSolverSettings):
        self.view = view            # keeps the device matrix alive (the C object only borrows it)
        self.settings = settings
        self._h = C.c_void_p()
        self._maxits = 1
        cs = settings.to_c()
        rc = lib.b200_prec_create(C.byref(cs), view._h, C.byref(self._h))
        if rc:
            raise ValueError(_lib.last_error())     # std::invalid_argument in the reference

    def dim(self) -> int:
        return lib.b200_prec_dim(self._h)

    def relaxationAvailable(self) -> bool:
        return bool(lib.b200_prec_relaxation_available(self._h))

    def setApplyParams(self, rtol=0.0, atol=0.0, dtol=0.0, ctol=False, maxits=1) -> None:
        """SolveParams (include/solverops_base.hpp:18-26); only maxits is used with ctol=False,
        which is what the PETSc glue sets (src/blasted_petsc.cpp:532)."""
        self._maxits = int(maxits)
        lib.b200_prec_set_apply_params(self._h, float(rtol), float(atol), float(dtol), int(ctol),
                                       int(maxits))

    def compute(self, vals=None) -> PrecInfo:
        """Preconditioner::compute().  With `vals` (host array of new matrix values, same pattern):
        the values are uploaded chunk by chunk with the layout conversion and the initial guess of
        the factor behind the copies (b200_prec_compute_host)."""
        info = np.zeros(6)
        if vals is None:
            check(lib.b200_prec_compute(self._h, info.ctypes.data_as(C.c_void_p)))
        else:
            _, vp = _host(vals)
            check(lib.b200_prec_compute_host(self._h, vp, info.ctypes.data_as(C.c_void_p)))
        return PrecInfo(info)

    def apply(self, r, z=None):
        if _is_torch_cuda(r):
            import torch
            z = torch.empty_like(r) if z is None else z
            check(lib.b200_prec_apply(self._h, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr())))
            return z
        _, rp = _host(r)
        z = np.zeros(self.dim()) if z is None else z
        _, zp = _host(z, True)
        check(lib.b200_prec_apply_host(self._h, rp, zp))
        return z

    def apply_relax(self, b, x):
        if _is_torch_cuda(b):
            check(lib.b200_prec_apply_relax(self._h, C.c_void_p(b.data_ptr()),
                                            C.c_void_p(x.data_ptr()), self._maxits))
            return x
        _, bp = _host(b)
        _, xp = _host(x, True)
        check(lib.b200_prec_apply_relax_host(self._h, bp, xp, self._maxits))
        return x

    # ---- setup products / diagnostics (parity checks) ----
    def set_sweeps(self, nbuild: int, napply: int) -> None:
        lib.b200_prec_set_sweeps(self._h, nbuild, napply)

    def set_stream(self, stream_ptr: int) -> None:
        lib.b200_prec_set_stream(self._h, C.c_void_p(stream_ptr))

    def ilu_positions(self):
        npos = C.c_longlong()
        check(lib.b200_prec_positions_size(self._h, C.byref(npos)))
        nnzb = self.view.m.nnzb
        posptr = np.empty(nnzb + 1, dtype=np.int32)
        lowerp = np.empty(max(npos.value, 1), dtype=np.int32)
        upperp = np.empty(max(npos.value, 1), dtype=np.int32)
        check(lib.b200_prec_get_positions(self._h, posptr.ctypes.data_as(C.c_void_p),
                                          lowerp.ctypes.data_as(C.c_void_p),
                                          upperp.ctypes.data_as(C.c_void_p)))
        return posptr, lowerp[:npos.value], upperp[:npos.value]

    def pattern_stats(self) -> dict:
        """Work-list sizes {nlower, nupper, nuwork, npos_l, npos_u} without copying the lists."""
        out = (C.c_longlong*5)()
        check(lib.b200_prec_pattern_stats(self._h, out))
        return dict(zip(("nlower", "nupper", "nuwork", "npos_l", "npos_u"), [int(v) for v in out]))

    def nlevels(self) -> int:
        nl = C.c_int()
        check(lib.b200_prec_levels_size(self._h, C.byref(nl)))
        return nl.value

    def levels(self):
        nl = C.c_int()
        check(lib.b200_prec_levels_size(self._h, C.byref(nl)))
        ptr = np.empty(nl.value + 1, dtype=np.int32)
        rows = np.empty(lib.b200_mat_nbrows(self.view._h), dtype=np.int32)
        check(lib.b200_prec_get_levels(self._h, ptr.ctypes.data_as(C.c_void_p),
                                       rows.ctypes.data_as(C.c_void_p)))
        return ptr, rows

    def factor(self) -> np.ndarray:
        m = self.view.m
        out = np.empty(m.nnzb * m.bs * m.bs)
        check(lib.b200_prec_get_factor(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def dblocks(self) -> np.ndarray:
        m = self.view.m
        out = np.empty(m.nbrows * m.bs * m.bs)
        check(lib.b200_prec_get_dblocks(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def scale_vector(self) -> np.ndarray:
        out = np.empty(self.dim())
        check(lib.b200_prec_get_scale(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def ilu_residual(self) -> float:
        r = C.c_double()
        check(lib.b200_prec_ilu_residual(self._h, C.byref(r)))
        return r.value

    def last_times(self):
        c, a = C.c_double(), C.c_double()
        check(lib.b200_prec_last_times(self._h, C.byref(c), C.byref(a)))
        return c.value, a.value

    def close(self):
        if self._h:
            lib.b200_prec_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SRFactory:
    """src/solverfactory.cpp:34-228"""

    def solverTypeFromString(self, precstr: str) -> int:
        if precstr not in SOLVER_TYPES:
            raise ValueError("BLASTed: Preconditioner type not available!")
        return SOLVER_TYPES[precstr]

    def create_preconditioner(self, prec_matrix, settings: AsyncSolverSettings) -> Preconditioner:
        """prec_matrix: an SRMatrix (host arrays, copied to the device) or an SRMatrixView already
        on the device (shared, as PETSc shares the local block between operator and PC)."""
        view = prec_matrix if isinstance(prec_matrix, SRMatrixView) else SRMatrixView(prec_matrix)
        return Preconditioner(view, settings)


# ---- Krylov drivers (tests/solvers.hpp:29-135) ----

class _IterativeSolver:
    _name = ""

    def __init__(self, mat: SRMatrixView, precond: Preconditioner):
        self.A, self.prec = mat, precond
        self.tol, self.maxiter, self.restart = 1e-6, 1000, 30

    def setParams(self, toler: float, maxits: int) -> None:
        self.tol, self.maxiter = toler, maxits

    def solve(self, b, x) -> SolveInfo:
        ci = _lib.SolveInfo()
        ph = self.prec._h if self.prec is not None else None
        if _is_torch_cuda(b):
            check(lib.b200_solve(self._name.encode(), self.A._h, ph, C.c_void_p(b.data_ptr()),
                                 C.c_void_p(x.data_ptr()), self.tol, self.maxiter, self.restart,
                                 C.byref(ci)))
        else:
            _, bp = _host(b)
            _, xp = _host(x, True)
            check(lib.b200_solve_host(self._name.encode(), self.A._h, ph, bp, xp, self.tol,
                                      self.maxiter, self.restart, C.byref(ci)))
        return SolveInfo(bool(ci.converged), ci.iters, ci.resnorm, ci.bnorm, ci.device_ms*1e-3,
                         ci.prec_ms*1e-3)


class RichardsonSolver(_IterativeSolver):
    _name = "richardson"


class BiCGSTAB(_IterativeSolver):
    _name = "bicgstab"


class GCR(_IterativeSolver):
    _name = "gcr"

    def __init__(self, mat, precond, n_restart: int):
        super().__init__(mat, precond)
        self.restart = n_restart


class FGMRES(GCR):
    """Flexible GMRES(m) proper (PETSc's -ksp_type fgmres): same iterates as GCR in exact arithmetic
    (tests/solvers.hpp:108-110) at about half the orthogonalisation traffic."""
    _name = "fgmres"


# ---- per-kernel-class device timing (b200_profile_*) ----
KERNEL_CLASSES = ["factor_lower", "factor_upper", "factor_init", "diag_invert", "tri_lower",
                  "tri_upper", "spmv", "other"]
# Krylov drivers and the partitioned layer: BLAS-1 launches, halo pack, time the compute stream
# waited for the halo, all-reduce of the dot products (includes waiting for the slowest rank)
ALL_CLASSES = KERNEL_CLASSES + ["blas1", "halo_pack", "halo_wait", "allreduce"]


def profile_enable(on: bool = True) -> None:
    lib.b200_profile_enable(int(on))


def profile_reset() -> None:
    lib.b200_profile_reset()


def profile_get():
    """{class: (total ms, launches)} accumulated since the last reset."""
    ms = np.zeros(8)
    cnt = np.zeros(8, dtype=np.int64)
    check(lib.b200_profile_get(ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES)}


def profile_get_all():
    """profile_get() over every class (ALL_CLASSES)."""
    n = lib.b200_profile_classes()
    assert n == len(ALL_CLASSES)
    ms = np.zeros(n)
    cnt = np.zeros(n, dtype=np.int64)
    check(lib.b200_profile_get_n(n, ms.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
    return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(ALL_CLASSES)}
