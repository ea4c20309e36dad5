"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic: row partitioning, halo plans and
the halo exchange order, checked with the oracle SpMV against the unpartitioned product."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from blasted_b200 import matgen
from blasted_b200.dist import partition_rows, poisson3d_slab, halo_exchange_host, row_offsets
from oracle import orc
from util import case, SEED


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _matrix(name):
    if name == "poisson":
        return matgen.poisson3d(0, 7, dims=(5, 4, 9))
    if name == "bsr4":
        return matgen.block_stencil((9, 11), 4, SEED)
    if name == "bsr5":
        return matgen.block_stencil((4, 3, 7), 5, SEED)
    return case(name)


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _matrix(name)
        part = partition_rows(m, world)[rank]
        x = np.random.default_rng(SEED).standard_normal(m.dim)
        bs = m.bs
        xl = x[part.row_begin*bs:part.row_end*bs]
        halo = halo_exchange_host(part, xl)
        y = orc().spmv(part.diag, xl)
        if part.offd is not None:
            y = y + orc().spmv(part.offd, halo)
        want = orc().spmv(m, x)[part.row_begin*bs:part.row_end*bs]
        err = float(np.abs(y - want).max()/np.abs(want).max())
        # a global dot product by all-reduce, as the Krylov drivers do
        import torch
        t = torch.tensor([float(xl @ xl)], dtype=torch.float64)
        dist.all_reduce(t)
        derr = abs(t.item() - float(x @ x))/float(x @ x)
        out[rank] = (err, derr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["poisson", "bsr4", "bsr5", "2dcyl1_bsr4", "msc00726_csr"])
def test_partitioned_spmv_matches_global(world, name):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, derr = out[r]
        assert err < 1e-13 and derr < 1e-13, (r, err, derr)


def test_partition_covers_matrix_and_plans_agree():
    m = matgen.block_stencil((7, 9), 4, SEED)
    for world in (1, 2, 4, 5):
        parts = partition_rows(m, world)
        assert sum(p.diag.nbrows for p in parts) == m.nbrows
        nnz = sum(p.diag.nnzb + (p.offd.nnzb if p.offd is not None else 0) for p in parts)
        assert nnz == m.nnzb
        for p in parts:
            assert sum(p.recv_counts) == p.nhalo
            for k, q in enumerate(p.neigh):
                # what p receives from q is what q sends to p
                kq = parts[q].neigh.index(p.rank)
                assert parts[q].send_counts[kq] == p.recv_counts[k]
    # single part: no halo at all
    p = partition_rows(m, 1)[0]
    assert p.offd is None and p.nhalo == 0 and p.neigh == []


@pytest.mark.parametrize("world", [1, 2, 4])
def test_poisson_slab_builder_equals_general_partition(world):
    dims = (4, 3, 8)
    m = matgen.poisson3d(0, 7, dims=dims)
    plane = dims[0]*dims[1]
    offs = row_offsets(dims[2], world)*plane
    parts = partition_rows(m, world, offsets=offs)
    for r in range(world):
        a, b = parts[r], poisson3d_slab(0, r, world, dims=dims)
        assert (a.row_begin, a.row_end, a.nhalo, a.neigh) == (b.row_begin, b.row_end, b.nhalo, b.neigh)
        assert a.send_counts == b.send_counts and a.recv_counts == b.recv_counts
        assert np.array_equal(a.send_idx, b.send_idx)
        for x, y in ((a.diag, b.diag), (a.offd, b.offd)):
            if x is None:
                assert y is None
                continue
            assert np.array_equal(x.browptr, y.browptr) and np.array_equal(x.bcolind, y.bcolind)
            assert np.array_equal(x.vals, y.vals)
