"""GPU tests of the distributed operator and Krylov drivers (NCCL).  With one visible GPU the
communicator has a single rank (no neighbours); with two or more the test spawns two ranks."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import blasted_b200 as bb
    from blasted_b200 import matgen
    from blasted_b200.dist import Comm, DistMatrix, partition_rows, poisson3d_slab
    from blasted_b200.solverfactory import SOLVER_TYPES
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = Comm.from_torch_distributed()
        res = {}
        # (1) partitioned SpMV == global SpMV, scalar and block
        for name, m in (("poisson", matgen.poisson3d(0, 7, dims=(12, 10, 16))),
                        ("bsr4", matgen.block_stencil((24, 20), 4, 7))):
            part = partition_rows(m, world)[rank]
            A = DistMatrix(comm, part)
            x = np.random.default_rng(3).standard_normal(m.dim)
            bs = m.bs
            xl = torch.from_numpy(x[part.row_begin*bs:part.row_end*bs]).cuda()
            y = A.apply(xl).cpu().numpy()
            want = orc().spmv(m, x)[part.row_begin*bs:part.row_end*bs]
            res["spmv_" + name] = float(np.abs(y - want).max()/np.abs(want).max())
        # (2) distributed GCR + block-Jacobi async ILU(0) on a Poisson slab problem
        n = 24
        part = poisson3d_slab(n, rank, world)
        A = DistMatrix(comm, part)
        s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=1, nbuildsweeps=20, napplysweeps=40)
        prec = bb.SRFactory().create_preconditioner(A.diag, s)
        prec.compute()
        ones = torch.ones(part.diag.dim, dtype=torch.float64, device="cuda")
        b = A.apply(ones)
        xs = torch.zeros_like(b)
        info = A.solve("gcr", prec, b, xs, tol=1e-9, maxiter=500, restart=30)
        res["iters"] = info.iters
        res["relres"] = info.resnorm/info.bnorm
        res["xerr"] = float((xs - 1.0).abs().max().item())
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_distributed_spmv_and_solve():
    import torch
    import torch.multiprocessing as mp
    world = 2 if torch.cuda.device_count() >= 2 else 1
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        res = out[r]
        assert res["spmv_poisson"] < 1e-13 and res["spmv_bsr4"] < 1e-13, res
        assert res["relres"] < 1e-9 and res["xerr"] < 1e-6, res
    assert len({out[r]["iters"] for r in range(world)}) == 1      # same count on every rank


def test_device_poisson_generator_matches_host():
    """matgen_device assembles exactly the arrays of matgen.poisson3d (the 512^3 and 27-point
    256^3 bench operators are built with it), also when the rows are processed in ragged chunks;
    the device block-stencil generator follows the host recipe (pattern identical, blocks
    block-diagonally dominant)."""
    import torch
    from blasted_b200 import matgen
    from blasted_b200 import matgen_device as md
    for dims, st, chunk in (((7, 5, 9), 7, 100), ((16, 12, 10), 7, 333), ((5, 5, 3), 27, 8), ((9, 7, 6), 27, 1 << 21)):
        m = matgen.poisson3d(0, st, dims=dims)
        bp, bc, slot, offs = md.stencil_pattern_device(dims, st == 27, rows_per_chunk=chunk)
        assert np.array_equal(bp.cpu().numpy(), m.browptr)
        assert np.array_equal(bc.cpu().numpy(), m.bcolind)
        bp, bc, v = md.poisson3d_device(dims, st)
        assert np.array_equal(v.cpu().numpy(), m.vals)
    for dims, bs in (((12, 9), 4), ((6, 5, 4), 5)):
        m = matgen.block_stencil(dims, bs, 1)
        nb, bp, bc, v = md.block_stencil_device(dims, bs, 1)
        assert nb == m.nbrows and np.array_equal(bp.cpu().numpy(), m.browptr)
        assert np.array_equal(bc.cpu().numpy(), m.bcolind)
        blocks = v.cpu().numpy().reshape(-1, bs, bs).transpose(0, 2, 1)          # logical [r, c]
        rows = np.repeat(np.arange(nb), np.diff(m.browptr))
        off = np.abs(blocks).sum(axis=2).max(axis=1)
        off[m.diagind] = 0
        rowsum = np.bincount(rows, weights=off, minlength=nb)
        d = blocks[m.diagind]
        assert np.all(np.abs(d[:, np.arange(bs), np.arange(bs)]) > rowsum[:, None] + 0.7)


def test_slab_device_solve_matches_host_assembled():
    """One rank: the device-assembled slab operator gives the same FGMRES run as the host-assembled
    one, and the solve without a host synchronisation per iteration stops at the same iteration
    count as the per-iteration form would (columns past convergence are discarded)."""
    import torch
    import blasted_b200 as bb
    from blasted_b200.dist import Comm, DistMatrix, poisson3d_slab, poisson3d_slab_device
    from blasted_b200.solverfactory import SOLVER_TYPES
    n = 20
    comm = Comm.single()
    runs = []
    for dev in (False, True):
        if dev:
            part, view = poisson3d_slab_device(n, 0, 1)
            A = DistMatrix(comm, part, view)
        else:
            A = DistMatrix(comm, poisson3d_slab(n, 0, 1))
        s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["seqilu0"], bs=1, nbuildsweeps=1, napplysweeps=1)
        prec = bb.SRFactory().create_preconditioner(A.diag, s)
        prec.compute()
        b = A.apply(torch.ones(A.local_dim(), dtype=torch.float64, device="cuda"))
        x = torch.zeros_like(b)
        info = A.solve("fgmres", prec, b, x, tol=1e-10, maxiter=300, restart=30)
        assert info.converged and float((x - 1).abs().max()) < 1e-7
        runs.append(info.iters)
    assert runs[0] == runs[1]


def test_profile_classes_of_a_partitioned_solve():
    """b200_profile_get_n: the Krylov and exchange classes are counted next to the kernel classes
    (what bench.py's per-iteration breakdown reads); the first eight agree with b200_profile_get."""
    import torch
    import blasted_b200 as bb
    from blasted_b200 import solverfactory as sf
    from blasted_b200.dist import Comm, DistMatrix, poisson3d_slab_device
    part, view = poisson3d_slab_device(16, 0, 1)
    A = DistMatrix(Comm.single(), part, view)
    s = bb.AsyncSolverSettings(prectype=sf.SOLVER_TYPES["ilu0"], bs=1, nbuildsweeps=2, napplysweeps=2)
    prec = bb.SRFactory().create_preconditioner(A.diag, s)
    prec.compute()
    b = A.apply(torch.ones(A.local_dim(), dtype=torch.float64, device="cuda"))
    x = torch.zeros_like(b)
    sf.profile_reset()
    sf.profile_enable(True)
    try:
        info = A.solve("fgmres", prec, b, x, tol=1e-8, maxiter=200, restart=30)
    finally:
        sf.profile_enable(False)
    assert info.converged
    every, basic = sf.profile_get_all(), sf.profile_get()
    for k in sf.KERNEL_CLASSES:
        assert every[k] == basic[k]
    assert every["blas1"][1] >= 3*info.iters and every["blas1"][0] > 0
    assert every["spmv"][1] >= info.iters
    assert every["tri_lower"][1] >= info.iters            # two lower sweeps are ONE pass (first two fused)
    assert every["tri_upper"][1] >= 2*info.iters
    assert every["allreduce"][1] == 0 and every["halo_wait"][1] == 0     # one subdomain: no exchange
    sf.profile_reset()
    assert all(v == (0.0, 0) for v in sf.profile_get_all().values())


@pytest.mark.parametrize("mk", [lambda m: m.poisson3d(0, 7, dims=(9, 7, 11)),
                                lambda m: m.poisson3d(0, 27, dims=(6, 5, 8)),
                                lambda m: m.block_stencil((13, 11), 3, 5),
                                lambda m: m.block_stencil((12, 10), 4, 6),
                                lambda m: m.block_stencil((7, 6, 5), 5, 7),
                                lambda m: m.block_stencil((6, 5), 7, 8)])
@pytest.mark.parametrize("nparts", [2, 3, 5])
def test_partitioned_product_with_caller_filled_halo(mk, nparts):
    """Every rank's share of a P-way partitioned product on ONE GPU: the halo is filled by the test
    from the plan's send lists (which also checks that send and receive orders agree), the coupling
    part runs over the boundary rows only (`rows_gemv_add_kernel`).  Reference product:
    src/kernels/matvecs.cpp:78-108 on the whole matrix."""
    import torch
    from blasted_b200 import matgen
    from blasted_b200.dist import Comm, DistMatrix, partition_rows
    from oracle import orc
    m = mk(matgen)
    bs = m.bs
    x = np.random.default_rng(11).standard_normal(m.dim)
    want = orc().spmv(m, x)
    parts = partition_rows(m, nparts)
    comm = Comm.single()
    for p in parts:
        # what each neighbour q sends to p, in p's receive order
        segs = []
        for q in p.neigh:
            pq = parts[q]
            k = pq.neigh.index(p.rank)
            o = int(np.sum(pq.send_counts[:k]))
            rows = pq.send_idx[o:o + pq.send_counts[k]].astype(np.int64) + pq.row_begin
            assert len(rows) == p.recv_counts[p.neigh.index(q)]
            segs.append((rows[:, None]*bs + np.arange(bs)[None, :]).reshape(-1))
        halo = x[np.concatenate(segs)] if segs else np.zeros(0)
        assert len(halo) == p.nhalo*bs
        A = DistMatrix(comm, p)
        xl = torch.from_numpy(x[p.row_begin*bs:p.row_end*bs].copy()).cuda()
        y = A.apply_with_halo(xl, torch.from_numpy(halo.copy()).cuda()).cpu().numpy()
        ref = want[p.row_begin*bs:p.row_end*bs]
        assert np.abs(y - ref).max() <= 1e-13*np.abs(want).max()
        A.close()
