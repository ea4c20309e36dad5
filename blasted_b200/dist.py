"""One subdomain per GPU: row partitioning, halo plans and the distributed operator / Krylov drivers.

BLASTed is the *local* preconditioner of a subdomain (README.md:4; include/blasted_petsc.h:4-6): under
PETSc's ``-pc_type bjacobi`` every rank hands it the sequential diagonal block of the distributed
matrix (src/blasted_petsc.cpp:229-238, 594-608) and PETSc does the distributed MatMult and VecDot.
This module is that outer layer for the device drivers: contiguous row blocks, the diagonal block in
local numbering for the preconditioner and the SpMV, the off-diagonal couplings against a halo
buffer filled by NCCL send/recv, Krylov dots by NCCL all-reduce (blasted_b200/csrc/dist.cu).

The partitioning itself is host logic (numpy) and is what the CPU tests exercise with gloo.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import lib, check
from .matgen import SRMatrix, find_diagind, poisson3d
from .solverfactory import SRMatrixView, Preconditioner, SolveInfo


@dataclass
class LocalPart:
    """What one rank holds of a row-partitioned matrix."""
    rank: int
    row_begin: int                     # first global (block) row owned
    row_end: int
    diag: SRMatrix                     # square diagonal block, local column numbering
    offd: Optional[SRMatrix]           # local rows x halo buffer; None if no couplings
    nhalo: int
    neigh: List[int] = field(default_factory=list)          # neighbour ranks, ascending
    send_counts: List[int] = field(default_factory=list)    # (block) entries sent to each neighbour
    send_idx: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))  # local rows, grouped
    recv_counts: List[int] = field(default_factory=list)    # halo segment length per neighbour


def row_offsets(nbrows: int, nparts: int) -> np.ndarray:
    """Contiguous, near-equal row blocks."""
    base, rem = divmod(nbrows, nparts)
    sizes = np.array([base + (1 if p < rem else 0) for p in range(nparts)], dtype=np.int64)
    return np.concatenate([[0], np.cumsum(sizes)])


def partition_rows(m: SRMatrix, nparts: int, offsets=None) -> List[LocalPart]:
    """Split a (block-)row matrix into `nparts` contiguous row blocks with their halo plans.

    The halo buffer of rank p is ordered by owner rank, then by global column; what rank q sends to
    p is exactly that list restricted to q's rows (so send and receive orders agree)."""
    offsets = row_offsets(m.nbrows, nparts) if offsets is None else np.asarray(offsets, dtype=np.int64)
    bs2 = m.bs*m.bs
    owner_of = lambda cols: np.searchsorted(offsets, cols, side="right") - 1
    needs = {}                                  # (p, q) -> sorted unique global cols p needs from q
    parts = []
    for p in range(nparts):
        r0, r1 = int(offsets[p]), int(offsets[p+1])
        e0, e1 = int(m.browptr[r0]), int(m.browptr[r1])
        cols = m.bcolind[e0:e1].astype(np.int64)
        rows = np.repeat(np.arange(r1 - r0), np.diff(m.browptr[r0:r1+1]))
        blocks = m.vals[e0*bs2:e1*bs2].reshape(-1, bs2)
        local = (cols >= r0) & (cols < r1)
        # diagonal block
        dptr = np.zeros(r1 - r0 + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows[local], minlength=r1 - r0), out=dptr[1:])
        dcol = (cols[local] - r0).astype(np.int32)
        diag = SRMatrix(r1 - r0, m.bs, dptr, dcol, np.ascontiguousarray(blocks[local]).reshape(-1),
                        find_diagind(dptr, dcol), m.rowmajor)
        # off-diagonal couplings
        ocols, orows, oblocks = cols[~local], rows[~local], blocks[~local]
        offd, nhalo, neigh, recv_counts = None, 0, [], []
        if len(ocols):
            own = owner_of(ocols)
            halo_cols = []
            for q in np.unique(own):
                need = np.unique(ocols[own == q])
                needs[(p, int(q))] = need
                neigh.append(int(q))
                recv_counts.append(len(need))
                halo_cols.append(need)
            halo_cols = np.concatenate(halo_cols)           # halo buffer order
            nhalo = len(halo_cols)
            # owners ascending and cols ascending within an owner => halo_cols is globally sorted
            hidx = np.searchsorted(halo_cols, ocols).astype(np.int32)
            optr = np.zeros(r1 - r0 + 1, dtype=np.int32)
            np.cumsum(np.bincount(orows, minlength=r1 - r0), out=optr[1:])
            offd = SRMatrix(r1 - r0, m.bs, optr, hidx, np.ascontiguousarray(oblocks).reshape(-1),
                            None, m.rowmajor)
        parts.append(LocalPart(p, r0, r1, diag, offd, nhalo, neigh, [], np.zeros(0, np.int32),
                               recv_counts))
    # send lists: q sends to p what p needs from q (pattern need not be symmetric)
    for q in range(nparts):
        targets = sorted(p for (p, qq) in needs if qq == q)
        part = parts[q]
        extra = [p for p in targets if p not in part.neigh]
        # a rank may have to send to someone it receives nothing from, and vice versa
        allneigh = sorted(set(part.neigh) | set(extra))
        recv = dict(zip(part.neigh, part.recv_counts))
        part.neigh = allneigh
        part.recv_counts = [recv.get(p, 0) for p in allneigh]
        sidx, scnt = [], []
        for p in allneigh:
            need = needs.get((p, q), np.zeros(0, np.int64))
            sidx.append((need - part.row_begin).astype(np.int32))
            scnt.append(len(need))
        part.send_counts = scnt
        part.send_idx = np.concatenate(sidx) if sidx else np.zeros(0, np.int32)
    return parts


def poisson3d_slab(n: int, rank: int, world: int, dims=None) -> LocalPart:
    """The rank's z-slab of the 7-point Laplacian on dims=(nx,ny,nz) (default n^3), built directly
    (no global matrix): identical to partition_rows(poisson3d(n), world, plane-aligned offsets)[rank]."""
    nx, ny, nz = (n, n, n) if dims is None else dims
    zoff = row_offsets(nz, world)
    z0, z1 = int(zoff[rank]), int(zoff[rank+1])
    return _slab_couplings(nx, ny, z0, z1, rank, world, poisson3d(0, 7, dims=(nx, ny, z1 - z0)))


def _slab_couplings(nx, ny, z0, z1, rank, world, diag) -> LocalPart:
    """Halo plan and off-diagonal block of the z-slab [z0, z1): one plane to each neighbour."""
    plane = nx*ny
    nloc = plane*(z1 - z0)
    neigh, send_counts, recv_counts, sidx = [], [], [], []
    orow, ocol = [], []
    hoff = 0
    if rank > 0:
        neigh.append(rank - 1); send_counts.append(plane); recv_counts.append(plane)
        sidx.append(np.arange(plane, dtype=np.int32))
        orow.append(np.arange(plane)); ocol.append(hoff + np.arange(plane)); hoff += plane
    if rank < world - 1:
        neigh.append(rank + 1); send_counts.append(plane); recv_counts.append(plane)
        sidx.append(np.arange(nloc - plane, nloc, dtype=np.int32))
        orow.append(np.arange(nloc - plane, nloc)); ocol.append(hoff + np.arange(plane)); hoff += plane
    offd = None
    if neigh:
        orow, ocol = np.concatenate(orow), np.concatenate(ocol)
        order = np.lexsort((ocol, orow))
        orow, ocol = orow[order], ocol[order]
        optr = np.zeros(nloc + 1, dtype=np.int32)
        np.cumsum(np.bincount(orow, minlength=nloc), out=optr[1:])
        offd = SRMatrix(nloc, 1, optr, ocol.astype(np.int32), np.full(len(ocol), -1.0), None)
    return LocalPart(rank, z0*plane, z1*plane, diag, offd, hoff, neigh, send_counts,
                     np.concatenate(sidx) if sidx else np.zeros(0, np.int32), recv_counts)


def poisson3d_slab_device(n: int, rank: int, world: int, dims=None):
    """poisson3d_slab with the diagonal block assembled on the device: returns (part, diag_view)
    where part.diag is None (no host copy of the big block) and the couplings to the neighbouring
    slabs (one plane each) are small host arrays as before."""
    nx, ny, nz = (n, n, n) if dims is None else dims
    zoff = row_offsets(nz, world)
    z0, z1 = int(zoff[rank]), int(zoff[rank+1])
    from .matgen_device import poisson3d_device
    browptr, bcolind, vals = poisson3d_device((nx, ny, z1 - z0))
    view = SRMatrixView.from_device(nx*ny*(z1 - z0), 1, browptr, bcolind, vals)
    del browptr, bcolind, vals
    part = _slab_couplings(nx, ny, z0, z1, rank, world, None)
    return part, view


# ------------------------------------------------------------------ host-side exchange (tests)

def halo_exchange_host(part: LocalPart, x_local: np.ndarray, group=None) -> np.ndarray:
    """Fill the halo buffer with torch.distributed point-to-point (any backend; CPU tensors with
    gloo).  Mirrors dist.cu::halo_exchange and is what the CPU tests run."""
    import torch
    import torch.distributed as dist
    bs = part.diag.bs
    halo = np.zeros(part.nhalo*bs)
    reqs, recv_bufs = [], []
    so = 0
    for k, q in enumerate(part.neigh):
        ns, nr = part.send_counts[k], part.recv_counts[k]
        if ns:
            idx = part.send_idx[so:so+ns]
            buf = torch.from_numpy(np.ascontiguousarray(x_local.reshape(-1, bs)[idx]).reshape(-1))
            reqs.append(dist.isend(buf, q, group=group))
        if nr:
            rb = torch.empty(nr*bs, dtype=torch.float64)
            reqs.append(dist.irecv(rb, q, group=group))
            recv_bufs.append((k, rb))
        so += ns
    for r in reqs:
        r.wait()
    ro = np.concatenate([[0], np.cumsum(part.recv_counts)])*bs
    for k, rb in recv_bufs:
        halo[ro[k]:ro[k+1]] = rb.numpy()
    return halo


# ------------------------------------------------------------------ device objects

def nccl_library_path() -> Optional[str]:
    """The libnccl.so.2 bundled with torch (the one its own collectives use)."""
    try:
        import torch
        cand = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib",
                            "libnccl.so.2")
        return cand if os.path.exists(cand) else None
    except Exception:
        return None


class Comm:
    """NCCL communicator of the library, bootstrapped over an initialised torch.distributed group."""

    def __init__(self, rank: int, world: int, unique_id: bytes):
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(lib.b200_comm_create(buf, rank, world, C.byref(self._h)))

    @staticmethod
    def from_torch_distributed() -> "Comm":
        import torch
        import torch.distributed as dist
        path = nccl_library_path()
        check(lib.b200_nccl_load(path.encode() if path else None))
        rank, world = dist.get_rank(), dist.get_world_size()
        idbuf = C.create_string_buffer(128)
        if rank == 0:
            check(lib.b200_comm_unique_id(idbuf))
        t = torch.tensor(list(idbuf.raw), dtype=torch.uint8)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        return Comm(rank, world, bytes(t.cpu().tolist()))

    @staticmethod
    def single() -> "Comm":
        """One subdomain: no NCCL is loaded or initialised."""
        return Comm(0, 1, bytes(128))

    def allreduce_sum(self, vals) -> np.ndarray:
        v = np.ascontiguousarray(vals, dtype=np.float64).copy()
        check(lib.b200_comm_allreduce_sum(self._h, v.ctypes.data_as(C.c_void_p), len(v)))
        return v

    def close(self):
        if self._h:
            lib.b200_comm_destroy(self._h)
            self._h = C.c_void_p()


class DistMatrix:
    """The rank's share of the partitioned operator on the device."""

    def __init__(self, comm: Comm, part: LocalPart, diag_view: Optional[SRMatrixView] = None):
        self.comm, self.part = comm, part
        self.diag = diag_view if diag_view is not None else SRMatrixView(part.diag)
        self.offd = SRMatrixView(part.offd) if part.offd is not None else None
        self._h = C.c_void_p()
        neigh = np.asarray(part.neigh, dtype=np.int32)
        sc = np.asarray(part.send_counts, dtype=np.int32)
        rc = np.asarray(part.recv_counts, dtype=np.int32)
        si = np.ascontiguousarray(part.send_idx, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p) if len(a) else None
        check(lib.b200_dist_mat_create(comm._h, self.diag._h, self.offd._h if self.offd else None,
                                       part.nhalo, len(neigh), p(neigh), p(sc), p(si), p(rc),
                                       C.byref(self._h)))

    def local_dim(self) -> int:
        return self.diag.dim()

    def apply(self, x, y=None):
        """y_local = (A x)_local; x, y torch CUDA tensors of the local length.  Collective."""
        import torch
        y = torch.empty_like(x) if y is None else y
        check(lib.b200_dist_mat_apply(self._h, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr())))
        return y

    def apply_with_halo(self, x, halo, y=None):
        """y_local = A_diag x + A_offd halo with a halo buffer the caller filled (plan order).  Not
        collective: no exchange is posted."""
        import torch
        y = torch.empty_like(x) if y is None else y
        hp = C.c_void_p(halo.data_ptr()) if halo is not None and halo.numel() else None
        check(lib.b200_dist_mat_apply_with_halo(self._h, C.c_void_p(x.data_ptr()), hp,
                                                C.c_void_p(y.data_ptr())))
        return y

    def solve(self, solver: str, prec: Optional[Preconditioner], b, x, tol=1e-8, maxiter=1000,
              restart=30) -> SolveInfo:
        """Distributed Krylov solve with the local (block-Jacobi) preconditioner.  Collective."""
        ci = _lib.SolveInfo()
        check(lib.b200_dist_solve(solver.encode(), self._h, prec._h if prec is not None else None,
                                  C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), tol, maxiter,
                                  restart, C.byref(ci)))
        return SolveInfo(bool(ci.converged), ci.iters, ci.resnorm, ci.bnorm, ci.device_ms*1e-3,
                         ci.prec_ms*1e-3)

    def close(self):
        if self._h:
            lib.b200_dist_mat_destroy(self._h)
            self._h = C.c_void_p()
