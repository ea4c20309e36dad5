"""Front end on the device: Matrix Market / coordinate input, reordering and scaling.

Host-side mirror of the reference's
  COOMatrix                        include/coomatrix.hpp:63-118, src/coomatrix.cpp:189-408
  getSRMatrixFromCOO               src/coomatrix.cpp:421-435
  Reordering / ReorderingScaling   include/reorderingscaling.hpp:42-133, src/reorderingscaling.cpp
over the C ABI of include/blasted_b200.h (b200_mat_create_coo, b200_mat_reorder, b200_mat_scale,
b200_vec_reorder, b200_vec_scale).  Parsing the text file is host work; the sort, the CSR/BSR
construction and every permutation / scaling run on the GPU (csrc/frontend.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib, check
from .solverfactory import SRMatrixView, _is_torch_cuda, COLMAJOR, ROWMAJOR

FORWARD, INVERSE = 0, 1          # RSApplyMode, include/reorderingscaling.hpp:29
ROW, COLUMN = 0, 1               # RSApplyDir,  include/reorderingscaling.hpp:31


class MatrixReadException(RuntimeError):
    """include/coomatrix.hpp:130-134"""


def _storage(order) -> int:
    if order == "rowmajor":
        return ROWMAJOR
    if order == "colmajor":
        return COLMAJOR
    raise RuntimeError("getSRMatrixFromCOO: invalid storage order!")       # src/coomatrix.cpp:433


class COOMatrix:
    """Coordinate matrix.  Entries are kept as read; ordering them is part of the device conversion."""

    def __init__(self):
        self.nrows = self.ncols = self.nnz = 0
        self.rowind = np.zeros(0, dtype=np.int32)
        self.colind = np.zeros(0, dtype=np.int32)
        self.values = np.zeros(0, dtype=np.float64)

    @classmethod
    def from_triplets(cls, nrows: int, rowind, colind, values, ncols: int | None = None) -> "COOMatrix":
        self = cls()
        self.nrows, self.ncols = int(nrows), int(nrows if ncols is None else ncols)
        self.rowind = np.ascontiguousarray(rowind, dtype=np.int32)
        self.colind = np.ascontiguousarray(colind, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        if not (len(self.rowind) == len(self.colind) == len(self.values)):
            raise ValueError("triplet arrays differ in length")
        self.nnz = len(self.values)
        return self

    def readMatrixMarket(self, file: str) -> None:
        """Reads a general, real/integer, coordinate Matrix Market file (src/coomatrix.cpp:189-221:
        array storage, pattern and symmetric files are refused with MatrixReadException)."""
        with open(file, "r") as f:
            banner = f.readline().split()
            if len(banner) < 5 or banner[0] != "%%MatrixMarket" or banner[1].lower() != "matrix":
                raise MatrixReadException("! COOMatrix: readMatrixMarket: not a Matrix Market matrix file.")
            storage, scalar, kind = (t.lower() for t in banner[2:5])
            if storage != "coordinate":
                raise MatrixReadException("! COOMatrix: readMatrixMarket: Can only read coordinate storage.")
            if scalar == "pattern":
                raise MatrixReadException("! COOMatrix: readMatrixMarket: Cannot read pattern matrices.")
            if kind != "general":
                raise MatrixReadException("! COOMatrix: readMatrixMarket: Can only read general matrices.")
            line = f.readline()
            while line.startswith("%") or not line.strip():
                line = f.readline()
            nrows, ncols, nnz = (int(t) for t in line.split()[:3])
            ncol_file = 4 if scalar == "complex" else 3
            if scalar == "complex":
                raise MatrixReadException("! COOMatrix: readMatrixMarket: Cannot read complex matrices.")
            body = np.loadtxt(f, dtype=np.float64, ndmin=2, max_rows=nnz, usecols=range(ncol_file))
        if body.shape[0] != nnz:
            raise MatrixReadException("! COOMatrix: readMatrixMarket: fewer entries than declared.")
        self.nrows, self.ncols, self.nnz = nrows, ncols, nnz
        self.rowind = np.ascontiguousarray(body[:, 0], dtype=np.int32) - 1
        self.colind = np.ascontiguousarray(body[:, 1], dtype=np.int32) - 1
        self.values = np.ascontiguousarray(body[:, 2], dtype=np.float64)

    def numrows(self) -> int:
        return self.nrows

    def numcols(self) -> int:
        return self.ncols

    def numnonzeros(self) -> int:
        return self.nnz

    def _convert(self, bs: int, storage: int) -> SRMatrixView:
        if self.nrows != self.ncols:
            raise ValueError("only square matrices can be converted")       # assert, coomatrix.cpp:306
        h = C.c_void_p()
        check(lib.b200_mat_create_coo(self.nrows, self.nnz, self.rowind.ctypes.data_as(C.c_void_p),
                                      self.colind.ctypes.data_as(C.c_void_p),
                                      self.values.ctypes.data_as(C.c_void_p), bs, storage, 0,
                                      C.byref(h)))
        return SRMatrixView.from_handle(h, bs, storage == ROWMAJOR)

    def convertToCSR(self) -> SRMatrixView:
        """src/coomatrix.cpp:262-297, result resident on the device"""
        return self._convert(1, COLMAJOR)

    def convertToBSR(self, bs: int, stor="colmajor") -> SRMatrixView:
        """src/coomatrix.cpp:299-403, result resident on the device"""
        return self._convert(bs, _storage(stor))


def getSRMatrixFromCOO(coo: COOMatrix, bs: int, block_storage_order: str = "colmajor") -> SRMatrixView:
    """src/coomatrix.cpp:421-435"""
    if bs == 1:
        return coo.convertToCSR()
    return coo.convertToBSR(bs, block_storage_order)


def _iptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Reordering:
    """Reordering<double,int,bs> with the ordering supplied by the caller
    (include/reorderingscaling.hpp:42-101; `compute` there is pure virtual - the orderings come from
    external packages such as MC64, which is outside the path)."""

    def __init__(self, bs: int = 1):
        self.bs = bs
        self.rp = self.cp = None

    def setOrdering(self, rord, cord, length: int | None = None) -> None:
        """src/reorderingscaling.cpp:51-75; either may be None"""
        def prep(o):
            if o is None:
                return None
            o = np.ascontiguousarray(o, dtype=np.int32)
            if length is not None:
                o = np.ascontiguousarray(o[:length])
            if len(o) and (np.sort(o) != np.arange(len(o))).any():
                raise ValueError("ordering is not a permutation")
            return o
        self.rp, self.cp = prep(rord), prep(cord)

    def isRowReordering(self) -> bool:
        return self.rp is not None and len(self.rp) > 0

    def applyOrdering(self, target, mode: int = FORWARD, direction: int | None = None):
        """Matrix (SRMatrixView; src/reorderingscaling.cpp:77-205) or vector (numpy array or CUDA
        tensor, with direction ROW/COLUMN; :211-266), in place."""
        if isinstance(target, SRMatrixView):
            if self.rp is None and self.cp is None:
                return target
            check(lib.b200_mat_reorder(target._h, _iptr(self.rp), _iptr(self.cp), int(mode), 0))
            _refetch(target)
            return target
        if direction is None:
            raise ValueError("a direction (ROW or COLUMN) is needed to reorder a vector")
        ord_ = self.rp if direction == ROW else self.cp
        if ord_ is None or len(ord_) == 0:
            return target
        return _vec_call(lib.b200_vec_reorder, target, len(ord_), self.bs, ord_, mode)


class ReorderingScaling(Reordering):
    """ReorderingScaling<double,int,bs> (include/reorderingscaling.hpp:106-133)"""

    def __init__(self, bs: int = 1):
        super().__init__(bs)
        self.rowscale = self.colscale = None

    def setScaling(self, rowscale, colscale) -> None:
        self.rowscale = None if rowscale is None else np.ascontiguousarray(rowscale, dtype=np.float64)
        self.colscale = None if colscale is None else np.ascontiguousarray(colscale, dtype=np.float64)

    def applyScaling(self, target, mode: int = FORWARD, direction: int | None = None):
        """Matrix (src/reorderingscaling.cpp:282-337) or vector (:340-368), in place."""
        if isinstance(target, SRMatrixView):
            if self.rowscale is None and self.colscale is None:
                return target
            check(lib.b200_mat_scale(target._h, _iptr(self.rowscale), _iptr(self.colscale), int(mode), 0))
            _refetch(target)
            return target
        if direction is None:
            raise ValueError("a direction (ROW or COLUMN) is needed to scale a vector")
        sc = self.rowscale if direction == ROW else self.colscale
        if sc is None or len(sc) == 0:
            return target
        return _vec_call(lib.b200_vec_scale, target, len(sc), self.bs, sc, mode)


def _refetch(view: SRMatrixView):
    """Host mirror of a view after the resident matrix changed."""
    bs, rowmajor = view.m.bs, view.m.rowmajor
    view._bs, view._rowmajor = bs, rowmajor
    view.m = None
    view.m = view.to_host()
    return view.m


def _vec_call(fn, vec, n, bs, arr, mode):
    if _is_torch_cuda(vec):
        import torch
        dt = torch.int32 if arr.dtype == np.int32 else torch.float64
        darr = torch.as_tensor(arr, dtype=dt, device=vec.device)
        if vec.dtype != torch.float64 or not vec.is_contiguous() or vec.numel() < n * bs:
            raise ValueError("vector must be a contiguous float64 tensor of n*bs entries")
        check(fn(C.c_void_p(vec.data_ptr()), n, bs, C.c_void_p(darr.data_ptr()), int(mode), 1))
        torch.cuda.synchronize()
        return vec
    if not (isinstance(vec, np.ndarray) and vec.dtype == np.float64 and vec.flags["C_CONTIGUOUS"]):
        raise ValueError("vector must be a contiguous float64 array")
    if vec.size < n * bs:
        raise ValueError("vector shorter than the ordering")
    check(fn(vec.ctypes.data_as(C.c_void_p), n, bs, arr.ctypes.data_as(C.c_void_p), int(mode), 0))
    return vec
