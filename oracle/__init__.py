"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE ONLY).

``orc``  : the plain-C restatement (oracle/blasted_oracle.c -> oracle/liboracle.so).
``ref()``: the unmodified reference, compiled from /root/reference into
           oracle/_ref/libblasted_ref.so (only where that tree, or a prebuilt .so, is present).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  Nothing under blasted_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libblasted_ref.so")
# the product's C++ adapters behind the reference's handles (tests only; loads libblasted_b200.so)
ADAPTER_SO = os.path.join(_HERE, "_ref", "libb200_adapters.so")

INIT_F = {"init_zero": 0, "init_original": 1, "init_sgs": 2, "init_none": 3}
INIT_A = {"init_zero": 0, "init_jacobi": 1, "init_none": 2}

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so, and libblasted_ref.so when the reference tree is present."""
    args = ["make", "-C", _HERE, "-j8", "all"]
    if force:
        args.insert(1, "-B")
    subprocess.run(args, check=True, capture_output=True)


def _opt(a):
    """Optional double array -> pointer or NULL."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class _Oracle:
    """Thin numpy-facing wrapper over liboracle.so (see blasted_oracle.h for citations)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = self.lib = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        L.orc_spmv.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_gemv3.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, C.c_double, _dp,
                                C.c_double, _dp, _dp]
        L.orc_ilu_positions.argtypes = [C.c_int, _ip, _ip, _ip, _ip, vp, vp]
        L.orc_ilu_positions.restype = C.c_longlong
        L.orc_compute_levels.argtypes = [C.c_int, _ip, _ip, _ip]
        L.orc_dag_levels.argtypes = [C.c_int, _ip, _ip, _ip, _ip]
        L.orc_scaling_vector.argtypes = [C.c_int, C.c_int, _dp, _ip, _dp]
        L.orc_block_inverse.argtypes = [C.c_int, C.c_int, _dp, _dp]
        L.orc_ilu0_init.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, vp, C.c_int, _dp]
        L.orc_ilu0_sweeps.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _ip, _ip,
                                      vp, C.c_int, _dp]
        L.orc_ilu0_sweep_synchronous.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip,
                                                 _ip, _ip, vp, _dp, _dp]
        L.orc_ilu0_invert_diag.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _dp]
        L.orc_ilu0_nonlinear_res.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _ip, _ip,
                                             _ip, vp, _dp]
        L.orc_ilu0_nonlinear_res.restype = C.c_double
        L.orc_matrix_abs_sum.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, vp]
        L.orc_matrix_abs_sum.restype = C.c_double
        L.orc_ilu0_apply.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _ip, _dp, vp, C.c_int,
                                     C.c_int, _dp, _dp, _dp]
        L.orc_jacobi_setup.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _dp]
        L.orc_jacobi_apply.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_sgs_apply.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _dp, C.c_int,
                                    C.c_int, _dp, _dp, _dp]
        L.orc_sgs_relax.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _dp, C.c_int,
                                    _dp, _dp]
        L.orc_gs_relax.argtypes = L.orc_sgs_relax.argtypes
        L.orc_jacobi_relax.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _dp, C.c_int,
                                       _dp, _dp, _dp]
        L.orc_diagonal_dominance.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp]
        L.orc_coo_convert.restype = C.c_longlong
        L.orc_coo_convert.argtypes = [C.c_int, C.c_longlong, _ip, _ip, _dp, C.c_int, C.c_int, vp, vp,
                                      vp, vp]
        L.orc_reorder_matrix.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, vp, vp, C.c_int]
        L.orc_reorder_vector.argtypes = [C.c_int, C.c_int, _ip, C.c_int, _dp]
        L.orc_scale_matrix.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, vp, vp, C.c_int]
        L.orc_scale_vector.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _dp]

    # ---- front end ----
    def coo_convert(self, nrows, rowind, colind, values, bs, rowmajor=False):
        """-> (browptr, bcolind, diagind, vals) as COOMatrix sort + convertToCSR/BSR produce them"""
        r = np.ascontiguousarray(rowind, dtype=np.int32)
        c = np.ascontiguousarray(colind, dtype=np.int32)
        v = np.ascontiguousarray(values, dtype=np.float64)
        nnzb = self.lib.orc_coo_convert(nrows, len(v), r, c, v, bs, int(rowmajor), None, None, None, None)
        nb = nrows // bs
        browptr = np.zeros(nb + 1, dtype=np.int32)
        bcolind = np.zeros(max(nnzb, 1), dtype=np.int32)
        diagind = np.zeros(max(nb, 1), dtype=np.int32)
        vals = np.zeros(max(nnzb * bs * bs, 1))
        self.lib.orc_coo_convert(nrows, len(v), r, c, v, bs, int(rowmajor), _opt(browptr),
                                 _opt(bcolind), _opt(diagind), _opt(vals))
        return browptr, bcolind[:nnzb], diagind[:nb], vals[:nnzb * bs * bs]

    def reorder_matrix(self, m, rord, cord, inverse=False):
        """-> (browptr, bcolind, vals) of the permuted matrix (copies)"""
        bp, bc, v = m.browptr.copy(), m.bcolind.copy(), m.vals.copy()
        ro = None if rord is None else np.ascontiguousarray(rord, dtype=np.int32)
        co = None if cord is None else np.ascontiguousarray(cord, dtype=np.int32)
        self.lib.orc_reorder_matrix(m.bs, m.nbrows, bp, bc, v, _opt(ro), _opt(co), int(inverse))
        return bp, bc, v

    def reorder_vector(self, bs, ordering, vec, inverse=False):
        o = np.ascontiguousarray(ordering, dtype=np.int32)
        out = np.array(vec, dtype=np.float64)
        self.lib.orc_reorder_vector(bs, len(o), o, int(inverse), out)
        return out

    def scale_matrix(self, m, rowscale, colscale, inverse=False):
        v = m.vals.copy()
        rs = None if rowscale is None else np.ascontiguousarray(rowscale, dtype=np.float64)
        cs = None if colscale is None else np.ascontiguousarray(colscale, dtype=np.float64)
        self.lib.orc_scale_matrix(m.bs, m.nbrows, m.browptr, m.bcolind, v, _opt(rs), _opt(cs),
                                  int(inverse))
        return v

    def scale_vector(self, bs, scale, vec, inverse=False):
        sc = np.ascontiguousarray(scale, dtype=np.float64)
        out = np.array(vec, dtype=np.float64)
        self.lib.orc_scale_vector(bs, len(sc), sc, int(inverse), out)
        return out

    # ---- matrix ops ----
    def spmv(self, m, x):
        y = np.empty(m.dim)
        self.lib.orc_spmv(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                          np.ascontiguousarray(x, dtype=np.float64), y)
        return y

    def gemv3(self, m, a, x, b, y):
        z = np.empty(m.dim)
        self.lib.orc_gemv3(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals, a,
                           np.ascontiguousarray(x), b, np.ascontiguousarray(y), z)
        return z

    # ---- setup products ----
    def ilu_positions(self, m):
        posptr = np.empty(m.nnzb + 1, dtype=np.int32)
        npos = self.lib.orc_ilu_positions(m.nbrows, m.browptr, m.bcolind, m.diagind, posptr,
                                          None, None)
        lowerp = np.empty(max(npos, 1), dtype=np.int32)
        upperp = np.empty(max(npos, 1), dtype=np.int32)
        self.lib.orc_ilu_positions(m.nbrows, m.browptr, m.bcolind, m.diagind, posptr,
                                   lowerp.ctypes.data_as(C.c_void_p),
                                   upperp.ctypes.data_as(C.c_void_p))
        return posptr, lowerp[:npos], upperp[:npos]

    def compute_levels(self, m):
        lv = np.empty(m.nbrows + 1, dtype=np.int32)
        n = self.lib.orc_compute_levels(m.nbrows, m.browptr, m.bcolind, lv)
        if n < 0:
            raise RuntimeError("Faulty dependency list!")
        return lv[:n].copy()

    def dag_levels(self, m):
        lv = np.empty(m.nbrows, dtype=np.int32)
        n = self.lib.orc_dag_levels(m.nbrows, m.browptr, m.bcolind, m.diagind, lv)
        return n, lv

    def scaling_vector(self, m):
        s = np.empty(m.dim)
        self.lib.orc_scaling_vector(m.bs, m.nbrows, m.vals, m.diagind, s)
        return s

    def block_inverse(self, bs, rowmajor, a):
        out = np.empty(bs * bs)
        self.lib.orc_block_inverse(bs, int(rowmajor), np.ascontiguousarray(a, dtype=np.float64), out)
        return out

    # ---- ILU(0) ----
    def ilu0_init(self, m, scale, fact_init):
        ilu = np.zeros(m.nnzb * m.bs * m.bs)
        fi = INIT_F[fact_init] if isinstance(fact_init, str) else fact_init
        self.lib.orc_ilu0_init(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                               m.diagind, _opt(scale), fi, ilu)
        return ilu

    def ilu0_sweeps(self, m, plist, scale, nsweeps, ilu):
        posptr, lowerp, upperp = plist
        lp = lowerp if len(lowerp) else np.zeros(1, np.int32)
        up = upperp if len(upperp) else np.zeros(1, np.int32)
        self.lib.orc_ilu0_sweeps(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                                 m.diagind, posptr, lp, up, _opt(scale), nsweeps, ilu)
        return ilu

    def ilu0_sweep_synchronous(self, m, plist, scale, ilu_old):
        posptr, lowerp, upperp = plist
        lp = lowerp if len(lowerp) else np.zeros(1, np.int32)
        up = upperp if len(upperp) else np.zeros(1, np.int32)
        new = np.empty_like(ilu_old)
        self.lib.orc_ilu0_sweep_synchronous(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind,
                                            m.vals, m.diagind, posptr, lp, up, _opt(scale),
                                            ilu_old, new)
        return new

    def ilu0_invert_diag(self, m, ilu):
        self.lib.orc_ilu0_invert_diag(m.bs, int(m.rowmajor), m.nbrows, m.diagind, ilu)
        return ilu

    def ilu0_nonlinear_res(self, m, plist, scale, ilu):
        posptr, lowerp, upperp = plist
        lp = lowerp if len(lowerp) else np.zeros(1, np.int32)
        up = upperp if len(upperp) else np.zeros(1, np.int32)
        return self.lib.orc_ilu0_nonlinear_res(m.bs, int(m.rowmajor), m.nbrows, m.browptr,
                                               m.bcolind, m.vals, m.diagind, posptr, lp, up,
                                               _opt(scale), ilu)

    def matrix_abs_sum(self, m, scale=None):
        return self.lib.orc_matrix_abs_sum(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind,
                                           m.vals, _opt(scale))

    def exact_ilu0(self, m, scale=None, invert_diag=False):
        """Exact ILU(0) = one sequential sweep from the original matrix
        (tests/solverops/async_ilu_convergence.cpp:462-490)."""
        plist = self.ilu_positions(m)
        ilu = self.ilu0_init(m, scale, "init_original")
        self.ilu0_sweeps(m, plist, scale, 1, ilu)
        if invert_diag:
            self.ilu0_invert_diag(m, ilu)
        return ilu

    def ilu0_apply(self, m, ilu, scale, napplysweeps, apply_init, r):
        z = np.empty(m.dim)
        y = np.zeros(m.dim)
        ai = INIT_A[apply_init] if isinstance(apply_init, str) else apply_init
        rc = self.lib.orc_ilu0_apply(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind,
                                     m.diagind, ilu, _opt(scale), napplysweeps, ai,
                                     np.ascontiguousarray(r, dtype=np.float64), z, y)
        if rc:
            raise RuntimeError(" scalar_ilu0_apply: Invalid init type!")
        return z

    # ---- Jacobi / SGS ----
    def jacobi_setup(self, m):
        d = np.empty(m.nbrows * m.bs * m.bs)
        self.lib.orc_jacobi_setup(m.bs, int(m.rowmajor), m.nbrows, m.vals, m.diagind, d)
        return d

    def jacobi_apply(self, m, dblocks, r):
        z = np.empty(m.dim)
        self.lib.orc_jacobi_apply(m.bs, int(m.rowmajor), m.nbrows, dblocks,
                                  np.ascontiguousarray(r, dtype=np.float64), z)
        return z

    def sgs_apply(self, m, dblocks, napplysweeps, apply_init, r):
        z = np.zeros(m.dim)
        y = np.zeros(m.dim)
        ai = INIT_A[apply_init] if isinstance(apply_init, str) else apply_init
        self.lib.orc_sgs_apply(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                               m.diagind, dblocks, napplysweeps, ai,
                               np.ascontiguousarray(r, dtype=np.float64), z, y)
        return z

    def sgs_relax(self, m, dblocks, maxits, b, x0):
        x = np.array(x0, dtype=np.float64, copy=True)
        self.lib.orc_sgs_relax(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                               m.diagind, dblocks, maxits, np.ascontiguousarray(b), x)
        return x

    def gs_relax(self, m, dblocks, nsweeps, b, x0):
        x = np.array(x0, dtype=np.float64, copy=True)
        self.lib.orc_gs_relax(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                              m.diagind, dblocks, nsweeps, np.ascontiguousarray(b), x)
        return x

    def jacobi_relax(self, m, dblocks, maxits, b, x0):
        x = np.array(x0, dtype=np.float64, copy=True)
        xt = np.empty_like(x)
        self.lib.orc_jacobi_relax(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                                  m.diagind, dblocks, maxits, np.ascontiguousarray(b), x, xt)
        return x

    def diagonal_dominance(self, m, vals):
        out = np.empty(4)
        self.lib.orc_diagonal_dominance(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.diagind,
                                        vals, out)
        return out


_orc = None


def orc() -> _Oracle:
    global _orc
    if _orc is None:
        _orc = _Oracle()
    return _orc


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class RefPrec:
    """A reference preconditioner object created through the reference's own SRFactory."""

    def __init__(self, lib, m, prectype, scale=False, nbuildsweeps=1, napplysweeps=1,
                 fact_init="init_original", apply_init="init_jacobi", thread_chunk_size=128,
                 compute_precinfo=False, through_b200_factory=False):
        self.lib, self.m = lib, m
        create = through_b200_factory if through_b200_factory else lib.ref_prec_create
        self.h = create(prectype.encode(), m.bs, int(m.rowmajor), int(scale),
                                     nbuildsweeps, napplysweeps, INIT_F[fact_init],
                                     INIT_A[apply_init], thread_chunk_size, int(compute_precinfo),
                                     m.nbrows, m.browptr, m.bcolind, m.vals, m.diagind)
        if not self.h:
            raise ValueError(lib.ref_last_error().decode())

    def compute(self):
        info = np.zeros(6)
        if self.lib.ref_prec_compute(self.h, info):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return info

    def apply(self, r):
        z = np.zeros(self.m.dim)
        if self.lib.ref_prec_apply(self.h, np.ascontiguousarray(r, dtype=np.float64), z):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return z

    def apply_relax(self, b, x0, maxits):
        x = np.array(x0, dtype=np.float64, copy=True)
        if self.lib.ref_prec_apply_relax(self.h, np.ascontiguousarray(b, dtype=np.float64), x, maxits):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return x

    def factor(self):
        n = self.m.nnzb * self.m.bs ** 2
        out = np.empty(n)
        if self.lib.ref_prec_get_factor(self.h, int(self.m.rowmajor), n, out):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out

    def dblocks(self):
        n = self.m.nbrows * self.m.bs ** 2
        out = np.empty(n)
        if self.lib.ref_prec_get_dblocks(self.h, int(self.m.rowmajor), n, out):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out

    def dim(self):
        return self.lib.ref_prec_dim(self.h)

    def close(self):
        if self.h:
            self.lib.ref_prec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Ref:
    """Wrapper over the unmodified reference build (oracle/_ref/libblasted_ref.so)."""

    def __init__(self):
        # RTLD_GLOBAL: libb200_adapters.so (loaded on demand) resolves the reference's classes here
        L = self.lib = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        self._adapters = None
        vp = C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_prec_create.restype = vp
        L.ref_prec_create.argtypes = [C.c_char_p] + [C.c_int] * 10 + [_ip, _ip, _dp, _ip]
        L.ref_prec_compute.argtypes = [vp, _dp]
        L.ref_prec_apply.argtypes = [vp, _dp, _dp]
        L.ref_prec_apply_relax.argtypes = [vp, _dp, _dp, C.c_int]
        L.ref_prec_dim.argtypes = [vp]
        L.ref_prec_get_factor.argtypes = [vp, C.c_int, C.c_longlong, _dp]
        L.ref_prec_get_dblocks.argtypes = [vp, C.c_int, C.c_longlong, _dp]
        L.ref_prec_destroy.argtypes = [vp]
        L.ref_spmv.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _dp, _dp]
        L.ref_gemv3.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, C.c_double, _dp,
                                C.c_double, _dp, _dp]
        L.ref_ilu_positions_create.restype = vp
        L.ref_ilu_positions_create.argtypes = [C.c_int, _ip, _ip, _ip]
        L.ref_ilu_positions_size.restype = C.c_longlong
        L.ref_ilu_positions_size.argtypes = [vp]
        L.ref_ilu_positions_copy.argtypes = [vp, _ip, _ip, _ip]
        L.ref_ilu_positions_destroy.argtypes = [vp]
        L.ref_compute_levels.argtypes = [C.c_int, _ip, _ip, _ip, _ip, C.c_int]
        L.ref_ilu_nonlinear_res.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _ip, vp, _dp, C.c_int,
                                            C.POINTER(C.c_double)]
        L.ref_solve.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _ip, _dp,
                                _dp, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int),
                                C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.ref_set_num_threads.argtypes = [C.c_int]
        L.ref_read_mtx.restype = vp
        L.ref_read_mtx.argtypes = [C.c_char_p, C.c_int, C.c_int]
        L.ref_srmat_nbrows.argtypes = [vp]
        L.ref_srmat_nnzb.argtypes = [vp]
        L.ref_srmat_copy.argtypes = [vp, _ip, _ip, _ip, _dp]
        L.ref_srmat_destroy.argtypes = [vp]
        L.ref_reorder_scale.argtypes = [C.c_int, C.c_int] + [vp] * 8 + [C.c_int, vp, vp]

    def adapters(self):
        """oracle/_ref/libb200_adapters.so: the product's C++ adapters (blasted_b200/host) behind
        the handles of this driver.  Loaded only when a test asks for it - the reference arm of
        bench.py never does, so that process maps nothing of the product."""
        if self._adapters is None:
            if not os.path.exists(ADAPTER_SO):
                raise RuntimeError("oracle/_ref/libb200_adapters.so not built")
            A = C.CDLL(ADAPTER_SO, mode=C.RTLD_GLOBAL)
            A.ref_prec_create_b200.restype = C.c_void_p
            A.ref_prec_create_b200.argtypes = self.lib.ref_prec_create.argtypes
            A.ref_reorder_scale_b200.argtypes = self.lib.ref_reorder_scale.argtypes
            self._adapters = A
        return self._adapters

    # ---- front end ----
    def read_mtx(self, path, bs, rowmajor=False):
        """COOMatrix::readMatrixMarket + getSRMatrixFromCOO -> (browptr, bcolind, diagind, vals)"""
        h = self.lib.ref_read_mtx(str(path).encode(), bs, int(rowmajor))
        if not h:
            raise RuntimeError(self.lib.ref_last_error().decode())
        nb, nnzb = self.lib.ref_srmat_nbrows(h), self.lib.ref_srmat_nnzb(h)
        browptr = np.zeros(nb + 1, dtype=np.int32)
        bcolind = np.zeros(max(nnzb, 1), dtype=np.int32)
        diagind = np.zeros(max(nb, 1), dtype=np.int32)
        vals = np.zeros(max(nnzb * bs * bs, 1))
        self.lib.ref_srmat_copy(h, browptr, bcolind, diagind, vals)
        self.lib.ref_srmat_destroy(h)
        return browptr, bcolind[:nnzb], diagind[:nb], vals[:nnzb * bs * bs]

    def reorder_scale(self, m, rord=None, cord=None, rowscale=None, colscale=None, inverse=False,
                      rowvec=None, colvec=None, through_b200=False):
        """Reordering / ReorderingScaling applied to copies -> (browptr, bcolind, vals, rowvec, colvec);
        through_b200: the product's device adapter behind the same reference interface."""
        def ia(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.int32)

        def da(a):
            return None if a is None else np.array(a, dtype=np.float64)
        bp, bc, v, di = m.browptr.copy(), m.bcolind.copy(), m.vals.copy(), m.diagind.copy()
        ro, co, rs, cs, rv, cv = ia(rord), ia(cord), da(rowscale), da(colscale), da(rowvec), da(colvec)
        fn = self.adapters().ref_reorder_scale_b200 if through_b200 else self.lib.ref_reorder_scale
        rc = fn(m.bs, m.nbrows, _opt(bp), _opt(bc), _opt(v), _opt(di), _opt(ro),
                                        _opt(co), _opt(rs), _opt(cs), int(inverse), _opt(rv), _opt(cv))
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return bp, bc, v, rv, cv

    def num_threads(self):
        return self.lib.ref_num_threads()

    def set_num_threads(self, n):
        self.lib.ref_set_num_threads(n)

    def prec(self, m, prectype, **kw) -> RefPrec:
        return RefPrec(self.lib, m, prectype, **kw)

    def prec_b200(self, m, prectype, **kw) -> RefPrec:
        """The product's device preconditioner created through its C++ B200Factory adapter
        (blasted_b200/host), i.e. behind the reference's own SRPreconditioner interface."""
        return RefPrec(self.lib, m, prectype, through_b200_factory=self.adapters().ref_prec_create_b200, **kw)

    def spmv(self, m, x):
        y = np.empty(m.dim)
        if self.lib.ref_spmv(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                             m.diagind, np.ascontiguousarray(x, dtype=np.float64), y):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return y

    def gemv3(self, m, a, x, b, y):
        z = np.empty(m.dim)
        if self.lib.ref_gemv3(m.bs, int(m.rowmajor), m.nbrows, m.browptr, m.bcolind, m.vals,
                              m.diagind, a, np.ascontiguousarray(x), b, np.ascontiguousarray(y), z):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return z

    def ilu_positions(self, m):
        h = self.lib.ref_ilu_positions_create(m.nbrows, m.browptr, m.bcolind, m.diagind)
        npos = self.lib.ref_ilu_positions_size(h)
        posptr = np.empty(m.nnzb + 1, dtype=np.int32)
        lowerp = np.empty(max(npos, 1), dtype=np.int32)
        upperp = np.empty(max(npos, 1), dtype=np.int32)
        self.lib.ref_ilu_positions_copy(h, posptr, lowerp, upperp)
        self.lib.ref_ilu_positions_destroy(h)
        return posptr, lowerp[:npos], upperp[:npos]

    def compute_levels(self, m):
        lv = np.empty(m.nbrows + 1, dtype=np.int32)
        n = self.lib.ref_compute_levels(m.nbrows, m.browptr, m.bcolind, m.diagind, lv, len(lv))
        if n < 0:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return lv[:n].copy()

    def ilu_nonlinear_res(self, m, scale, ilu, chunk=128):
        out = C.c_double()
        if self.lib.ref_ilu_nonlinear_res(m.bs, m.nbrows, m.browptr, m.bcolind, m.vals, m.diagind,
                                          _opt(scale), ilu, chunk, C.byref(out)):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out.value

    def solve(self, solver, prec: RefPrec, m, b, x0=None, tol=1e-10, maxiter=1000, restart=30):
        x = np.zeros(m.dim) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        its, rr, wt = C.c_int(), C.c_double(), C.c_double()
        if self.lib.ref_solve(solver.encode(), prec.h, m.bs, int(m.rowmajor), m.nbrows, m.browptr,
                              m.bcolind, m.vals, m.diagind, np.ascontiguousarray(b), x, tol,
                              maxiter, restart, C.byref(its), C.byref(rr), C.byref(wt)):
            raise RuntimeError(self.lib.ref_last_error().decode())
        return x, its.value, rr.value, wt.value


_ref = None


def ref() -> _Ref:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libblasted_ref.so not built (reference tree absent)")
        _ref = _Ref()
    return _ref
