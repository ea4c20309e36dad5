"""GPU: partition-matched Krylov iteration counts (SURVEY.md 8(e) caveat).

Across P GPUs the preconditioner is block-Jacobi: one local BLASTed preconditioner per diagonal block
of the row partition, couplings to other subdomains dropped from M, exactly what PETSc's
`-pc_type bjacobi -sub_pc_type shell` hands the reference (src/blasted_petsc.cpp:594-608, one block per
rank).  Iteration counts therefore have to be judged against the REFERENCE run with the same
partition: here the reference's own preconditioner (oracle/_ref, unmodified sources) is built on
the block-diagonal part of the matrix for P = 2, 4, 8 and driven by the reference's own GCR
(tests/solvers.cpp:252-352) on the full operator - a host-side partitioned Krylov loop - and the
device solve with the same partition must need the same number of iterations within 5 %.
On one GPU the P subdomains are emulated the same way (preconditioner built on the block-diagonal
matrix, operator = full matrix); with two or more GPUs tests/test_gpu_dist.py runs real ranks."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.dist import row_offsets
from blasted_b200.solverfactory import SOLVER_TYPES
from oracle import have_ref, ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs the reference tree)")]


def block_diagonal_part(m, offsets):
    """The matrix with every coupling between different row blocks removed."""
    rows = np.repeat(np.arange(m.nbrows), np.diff(m.browptr))
    part_of = lambda idx: np.searchsorted(offsets, idx, side="right") - 1
    keep = part_of(rows) == part_of(m.bcolind)
    ptr = np.zeros(m.nbrows + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows[keep], minlength=m.nbrows), out=ptr[1:])
    bs2 = m.bs*m.bs
    vals = m.vals.reshape(-1, bs2)[keep].reshape(-1)
    col = m.bcolind[keep]
    return matgen.SRMatrix(m.nbrows, m.bs, ptr, col, np.ascontiguousarray(vals),
                           matgen.find_diagind(ptr, col), m.rowmajor)


def device_its(m, mbj, ptype, solver, b, tol, **kw):
    A, Abj = bb.SRMatrixView(m), bb.SRMatrixView(mbj)
    p = bb.SRFactory().create_preconditioner(Abj, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES[ptype], bs=m.bs, **kw))
    p.compute()
    sol = (bb.GCR(A, p, 30) if solver == "gcr" else bb.FGMRES(A, p, 30))
    sol.setParams(tol, 2000)
    x = np.zeros(m.dim)
    info = sol.solve(b, x)
    assert info.resnorm/info.bnorm < tol*1.01
    return info.iters, x


@pytest.mark.parametrize("nparts", [1, 2, 4, 8])
def test_block_jacobi_iteration_counts_match_partitioned_reference(nparts):
    n = 32
    m = matgen.poisson3d(n)
    plane = n*n
    offsets = row_offsets(n, nparts)*plane                   # z-slabs, as bench.py partitions C5
    mbj = block_diagonal_part(m, offsets)
    b = ref().spmv(m, np.ones(m.dim))
    tol = 1e-8
    R = ref()
    rp = R.prec(mbj, "seqilu0", nbuildsweeps=1, napplysweeps=1)
    rp.compute()
    xr, its_ref, rr, _ = R.solve("gcr", rp, m, b, tol=tol, maxiter=2000, restart=30)
    rp.close()
    assert rr < tol*1.01
    slack = lambda want: max(1, int(np.ceil(0.05*want)))
    # exact local factor + exact local solves: the reference's sequential preconditioner
    its, x = device_its(m, mbj, "seqilu0", "gcr", b, tol, nbuildsweeps=1)
    assert abs(its - its_ref) <= slack(its_ref), (nparts, its, its_ref)
    assert np.abs(x - xr).max() < 1e-5
    # asynchronous local factor and sweeps, converged (32 planes: 94 dependency levels)
    its_a, _ = device_its(m, mbj, "ilu0", "gcr", b, tol, nbuildsweeps=30, napplysweeps=100)
    assert abs(its_a - its_ref) <= slack(its_ref), (nparts, its_a, its_ref)
    # FGMRES(30) proper builds the same iterates as GCR(30) in exact arithmetic (tests/solvers.hpp:108-110)
    its_f, _ = device_its(m, mbj, "seqilu0", "fgmres", b, tol, nbuildsweeps=1)
    assert abs(its_f - its_ref) <= slack(its_ref), (nparts, its_f, its_ref)


def test_block_jacobi_weakens_as_in_the_reference():
    """The growth of the iteration count with the number of subdomains (dropped couplings) is the
    reference's own: same counts for P = 1 and P = 8 within 5 %, and P = 8 needs more than P = 1."""
    n = 24
    m = matgen.poisson3d(n)
    b = ref().spmv(m, np.ones(m.dim))
    R = ref()
    counts = {}
    for nparts in (1, 8):
        mbj = block_diagonal_part(m, row_offsets(n, nparts)*n*n)
        rp = R.prec(mbj, "seqilu0", nbuildsweeps=1, napplysweeps=1)
        rp.compute()
        _, its_ref, _, _ = R.solve("gcr", rp, m, b, tol=1e-8, maxiter=2000, restart=30)
        rp.close()
        its, _ = device_its(m, mbj, "seqilu0", "gcr", b, 1e-8, nbuildsweeps=1)
        counts[nparts] = (its, its_ref)
        assert abs(its - its_ref) <= max(1, int(np.ceil(0.05*its_ref)))
    assert counts[8][1] > counts[1][1] and counts[8][0] > counts[1][0]
