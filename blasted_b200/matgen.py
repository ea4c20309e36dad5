"""Synthetic sparse (block-)row matrices of the shapes named in BASELINE.json / SURVEY.md section 8(d).

Host-side (numpy) generators used by the tests and by bench.py.  The layout is the reference's
``SRMatrixStorage`` (include/srmatrixdefs.hpp:38-79): ``browptr[nbrows+1]``, ``bcolind[nnzb]`` (sorted
ascending in every row), ``vals[nnzb*bs*bs]`` (blocks contiguous, column-major inside a block unless
``rowmajor``), ``diagind[nbrows]``; int32 indices, fp64 values.

The 7-point operator follows tests/poisson3d-fd/poisson3d_fd.cpp:108-139 on a uniform grid with
Dirichlet boundaries (neighbours outside the grid are dropped).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SRMatrix:
    """A square sparse (block-)row matrix in the reference's raw layout."""
    nbrows: int
    bs: int
    browptr: np.ndarray
    bcolind: np.ndarray
    vals: np.ndarray
    diagind: np.ndarray
    rowmajor: bool = False

    @property
    def nnzb(self) -> int:
        return int(self.browptr[-1])

    @property
    def dim(self) -> int:
        return self.nbrows * self.bs

    def to_scipy(self):
        """scipy.sparse CSR (bs == 1) or BSR matrix; test helper."""
        import scipy.sparse as sp
        if self.bs == 1:
            return sp.csr_matrix((self.vals, self.bcolind, self.browptr),
                                 shape=(self.nbrows, self.nbrows))
        blocks = self.vals.reshape(-1, self.bs, self.bs)
        if not self.rowmajor:
            blocks = blocks.transpose(0, 2, 1)
        return sp.bsr_matrix((blocks, self.bcolind, self.browptr),
                             shape=(self.dim, self.dim))


def find_diagind(browptr: np.ndarray, bcolind: np.ndarray, strict: bool = True) -> np.ndarray:
    """Position of the diagonal (block) of every row in bcolind (cf. Mat_SeqAIJ::diag).

    With strict=False rows without a stored diagonal get -1 (usable for SpMV only)."""
    nbrows = len(browptr) - 1
    rows = np.repeat(np.arange(nbrows, dtype=np.int64), np.diff(browptr))
    pos = np.nonzero(bcolind == rows)[0]
    if len(pos) != nbrows:
        if strict:
            raise ValueError("matrix has a structurally missing diagonal")
        out = np.full(nbrows, -1, dtype=np.int32)
        out[rows[pos]] = pos
        return out
    return pos.astype(np.int32)


def _stencil_pattern(dims, offsets):
    """Pattern of a structured stencil with lexicographic (x fastest) numbering.

    dims = (nx, ny[, nz]); offsets = list of integer tuples (dx, dy[, dz]) in ascending column
    order.  Returns (browptr, bcolind, slot) where slot[k] is the index into `offsets` of entry k.
    """
    nd = len(dims)
    n = int(np.prod(dims))
    idx = np.arange(n, dtype=np.int64)
    coords = []
    rem = idx
    for d in range(nd):
        coords.append(rem % dims[d])
        rem = rem // dims[d]
    strides = [1]
    for d in range(1, nd):
        strides.append(strides[-1] * dims[d - 1])
    nof = len(offsets)
    valid = np.ones((n, nof), dtype=bool)
    cols = np.empty((n, nof), dtype=np.int64)
    for s, off in enumerate(offsets):
        lin = 0
        ok = np.ones(n, dtype=bool)
        for d in range(nd):
            c = coords[d] + off[d]
            ok &= (c >= 0) & (c < dims[d])
            lin += off[d] * strides[d]
        valid[:, s] = ok
        cols[:, s] = idx + lin
    browptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=browptr[1:])
    if browptr[-1] >= 2**31:
        raise ValueError("pattern exceeds int32 indexing")
    bcolind = cols[valid].astype(np.int32)
    slot = np.broadcast_to(np.arange(nof, dtype=np.int8), (n, nof))[valid]
    return browptr.astype(np.int32), bcolind, slot


def _offsets(nd: int, full: bool):
    """Stencil offsets in ascending linear-index order: star (2*nd+1) or full (3**nd)."""
    import itertools
    offs = []
    for t in itertools.product((-1, 0, 1), repeat=nd):       # t = (d_{nd-1}, ..., d_0)
        off = tuple(reversed(t))                               # (dx, dy, dz)
        if full or sum(abs(o) for o in off) <= 1:
            offs.append(off)
    return offs


def poisson3d(n: int, stencil: int = 7, dims=None) -> SRMatrix:
    """Scalar FD Laplacian on an n^3 (or dims = (nx,ny,nz)) interior grid.

    stencil=7 : diag 6, off -1 (h = 1 scaling of poisson3d_fd.cpp:118-137 on a uniform grid);
    stencil=27: diag 26, off -1.
    """
    dims = tuple(dims) if dims is not None else (n, n, n)
    offs = _offsets(3, stencil == 27)
    browptr, bcolind, slot = _stencil_pattern(dims, offs)
    centre = offs.index((0, 0, 0))
    diagval = 6.0 if stencil == 7 else 26.0
    vals = np.where(slot == centre, diagval, -1.0).astype(np.float64)
    return SRMatrix(len(browptr) - 1, 1, browptr, bcolind, vals, find_diagind(browptr, bcolind))


def poisson2d(nx: int, ny: int) -> SRMatrix:
    offs = _offsets(2, False)
    browptr, bcolind, slot = _stencil_pattern((nx, ny), offs)
    centre = offs.index((0, 0))
    vals = np.where(slot == centre, 4.0, -1.0).astype(np.float64)
    return SRMatrix(len(browptr) - 1, 1, browptr, bcolind, vals, find_diagind(browptr, bcolind))


def block_stencil(dims, bs: int, seed: int, rowmajor: bool = False) -> SRMatrix:
    """Synthetic flow-Jacobian-like BSR matrix.

    Star stencil of bs x bs blocks on a structured grid of `dims` cells.  Off-diagonal blocks are
    U(-0.5,0.5)/4; the diagonal block is (sum_j ||A_ij||_inf + 1) I + U(-0.25,0.25): block
    diagonally dominant, non-symmetric values, symmetric pattern.
    """
    nd = len(dims)
    offs = _offsets(nd, False)
    browptr, bcolind, slot = _stencil_pattern(tuple(dims), offs)
    nbrows = len(browptr) - 1
    nnzb = int(browptr[-1])
    rng = np.random.default_rng(seed)
    blocks = rng.uniform(-0.5, 0.5, size=(nnzb, bs, bs)) / 4.0      # logical [r, c]
    centre = offs.index(tuple([0] * nd))
    isdiag = slot == centre
    # infinity norm of every block, summed over the off-diagonal blocks of a row
    binf = np.abs(blocks).sum(axis=2).max(axis=1)
    binf[isdiag] = 0.0
    rows = np.repeat(np.arange(nbrows), np.diff(browptr))
    rowsum = np.bincount(rows, weights=binf, minlength=nbrows)
    dpos = np.nonzero(isdiag)[0]
    dblk = rng.uniform(-0.25, 0.25, size=(nbrows, bs, bs))
    dblk[:, np.arange(bs), np.arange(bs)] += (rowsum + 1.0)[:, None]
    blocks[dpos] = dblk
    if not rowmajor:
        blocks = blocks.transpose(0, 2, 1)
    vals = np.ascontiguousarray(blocks).reshape(-1)
    return SRMatrix(nbrows, bs, browptr, bcolind, vals, dpos.astype(np.int32), rowmajor)


def csr_to_bsr(m: SRMatrix, bs: int, rowmajor: bool = False, strict_diag: bool = True) -> SRMatrix:
    """Re-block a scalar CSR matrix (dimension divisible by bs) into BSR; zero-fills blocks."""
    assert m.bs == 1 and m.nbrows % bs == 0
    a = m.to_scipy().tobsr(blocksize=(bs, bs))
    a.sort_indices()
    blocks = a.data
    if not rowmajor:
        blocks = blocks.transpose(0, 2, 1)
    browptr = a.indptr.astype(np.int32)
    bcolind = a.indices.astype(np.int32)
    return SRMatrix(m.nbrows // bs, bs, browptr, bcolind,
                    np.ascontiguousarray(blocks, dtype=np.float64).reshape(-1),
                    find_diagind(browptr, bcolind, strict_diag), rowmajor)


def from_scipy(a, bs: int = 1, rowmajor: bool = False, strict_diag: bool = True) -> SRMatrix:
    """Build an SRMatrix from any scipy sparse matrix (sorted, duplicates summed)."""
    import scipy.sparse as sp
    a = sp.csr_matrix(a)
    a.sum_duplicates()
    a.sort_indices()
    m = SRMatrix(a.shape[0], 1, a.indptr.astype(np.int32), a.indices.astype(np.int32),
                 a.data.astype(np.float64), None)
    m.diagind = find_diagind(m.browptr, m.bcolind, strict_diag)
    return m if bs == 1 else csr_to_bsr(m, bs, rowmajor, strict_diag)
