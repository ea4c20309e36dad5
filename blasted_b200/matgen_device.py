"""Synthetic matrices of the BASELINE.json shapes assembled ON the device (torch CUDA tensors).

The host generators of matgen.py (numpy) are what the parity tests use; at the full bench sizes
(7-point 512^3: 9.4e8 entries, 27-point 256^3: 4.5e8 entries, BSR5 128^3: 2.9 GB of blocks) building
on the host and uploading would take longer than everything that is measured, so bench.py and
tools/config_report.py assemble with these instead.  Same stencils and the same recipe as matgen.py;
the scalar operators are entry-for-entry identical to the host generators
(tests/test_gpu_dist.py::test_device_poisson_generator_matches_host), the block matrices use torch's
generator instead of numpy's (same distribution, different stream).

Reference for the 7-point operator: tests/poisson3d-fd/poisson3d_fd.cpp:108-139 (uniform grid,
Dirichlet boundaries: neighbours outside the grid are dropped).
"""
from __future__ import annotations

import itertools


def _stencil_offsets(nd: int, full: bool):
    offs = []
    for t in itertools.product((-1, 0, 1), repeat=nd):         # (d_{nd-1}, ..., d_0): ascending linear index
        off = tuple(reversed(t))                                 # (dx, dy[, dz])
        if full or sum(abs(o) for o in off) <= 1:
            offs.append(off)
    return offs


def stencil_pattern_device(dims, full: bool, device="cuda", rows_per_chunk: int = 1 << 21):
    """Pattern of a structured stencil with lexicographic (x fastest) numbering, built chunk by
    chunk.  Returns (browptr int32, bcolind int32, slot int8) with slot = index into the offsets."""
    import torch
    nd = len(dims)
    n = 1
    for d in dims:
        n *= d
    offs = _stencil_offsets(nd, full)
    if len(offs)*n >= 2**31:
        raise ValueError("pattern exceeds int32 indexing")
    strides = [1]
    for d in range(1, nd):
        strides.append(strides[-1]*dims[d-1])
    lin = torch.tensor([sum(o[d]*strides[d] for d in range(nd)) for o in offs], dtype=torch.int64, device=device)
    slots = torch.arange(len(offs), dtype=torch.int8, device=device)
    counts, cols, slot = [], [], []
    for r0 in range(0, n, rows_per_chunk):
        idx = torch.arange(r0, min(n, r0 + rows_per_chunk), dtype=torch.int64, device=device)
        coords, rem = [], idx
        for d in range(nd):
            coords.append(rem % dims[d])
            rem = rem // dims[d]
        valid = torch.ones((idx.numel(), len(offs)), dtype=torch.bool, device=device)
        for s, o in enumerate(offs):
            ok = valid[:, s]
            for d in range(nd):
                if o[d] < 0:
                    ok = ok & (coords[d] > 0)
                elif o[d] > 0:
                    ok = ok & (coords[d] < dims[d] - 1)
            valid[:, s] = ok
        counts.append(valid.sum(dim=1, dtype=torch.int32))
        cols.append((idx[:, None] + lin[None, :])[valid].to(torch.int32))
        slot.append(slots[None, :].expand(idx.numel(), -1)[valid])
    counts = torch.cat(counts)
    browptr = torch.zeros(n + 1, dtype=torch.int32, device=device)
    browptr[1:] = torch.cumsum(counts, 0, dtype=torch.int64).to(torch.int32)
    return browptr, torch.cat(cols), torch.cat(slot), offs


def poisson3d_device(dims, stencil: int = 7, device="cuda"):
    """matgen.poisson3d(dims=dims, stencil=...) on the device: (browptr, bcolind, vals)."""
    import torch
    browptr, bcolind, slot, offs = stencil_pattern_device(tuple(dims), stencil == 27, device)
    centre = offs.index((0, 0, 0))
    vals = torch.where(slot == centre, 6.0 if stencil == 7 else 26.0, -1.0).to(torch.float64)
    return browptr, bcolind, vals


def block_stencil_device(dims, bs: int, seed: int, device="cuda"):
    """matgen.block_stencil on the device (column-major blocks): star stencil of bs x bs blocks,
    off-diagonal blocks U(-0.5,0.5)/4, diagonal block (sum_j ||A_ij||_inf + 1) I + U(-0.25,0.25).
    Returns (nbrows, browptr, bcolind, vals)."""
    import torch
    nd = len(dims)
    browptr, bcolind, slot, offs = stencil_pattern_device(tuple(dims), False, device)
    nbrows = browptr.numel() - 1
    nnzb = bcolind.numel()
    gen = torch.Generator(device=device).manual_seed(seed)
    blocks = (torch.rand((nnzb, bs, bs), dtype=torch.float64, device=device, generator=gen) - 0.5)/4.0
    isdiag = slot == offs.index(tuple([0]*nd))
    binf = blocks.abs().sum(dim=2).max(dim=1).values
    binf[isdiag] = 0.0
    rows = torch.repeat_interleave(torch.arange(nbrows, device=device), (browptr[1:] - browptr[:-1]).long())
    rowsum = torch.zeros(nbrows, dtype=torch.float64, device=device).index_add_(0, rows, binf)
    dblk = torch.rand((nbrows, bs, bs), dtype=torch.float64, device=device, generator=gen)*0.5 - 0.25
    dblk += torch.eye(bs, dtype=torch.float64, device=device)[None]*(rowsum + 1.0)[:, None, None]
    blocks[isdiag] = dblk
    vals = blocks.transpose(1, 2).contiguous().reshape(-1)        # column-major inside a block
    return nbrows, browptr, bcolind, vals
