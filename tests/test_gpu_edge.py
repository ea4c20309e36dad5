"""GPU edge cases through the C ABI: empty and one-row matrices, a missing diagonal, single-block
matrices of every block size, rows far above the staged kernels' tile capacity."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES
from oracle import orc
from util import relerr

pytestmark = pytest.mark.gpu


def settings(name, bs, **kw):
    return bb.AsyncSolverSettings(prectype=SOLVER_TYPES[name], bs=bs, **kw)


def test_empty_matrix():
    m = matgen.SRMatrix(0, 1, np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32),
                        np.zeros(0), np.zeros(0, dtype=np.int32))
    A = bb.SRMatrixView(m)
    assert A.dim() == 0
    assert A.apply(np.zeros(0)).shape == (0,)
    for name in ("jacobi", "sgs", "ilu0", "seqilu0", "level_sgs", "async_level_ilu0"):
        p = bb.SRFactory().create_preconditioner(A, settings(name, 1))
        p.compute()
        assert p.apply(np.zeros(0)).shape == (0,)


@pytest.mark.parametrize("bs", [1, 4, 5])
def test_single_block_row(bs):
    rng = np.random.default_rng(bs)
    blk = rng.standard_normal((bs, bs)) + 4 * np.eye(bs)
    m = matgen.SRMatrix(1, bs, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32),
                        np.ascontiguousarray(blk.T).reshape(-1), np.array([0], dtype=np.int32))
    A = bb.SRMatrixView(m)
    r = rng.standard_normal(bs)
    assert relerr(A.apply(r), blk @ r) < 1e-13
    for name in ("jacobi", "sgs", "ilu0", "seqilu0", "level_sgs", "async_level_ilu0"):
        p = bb.SRFactory().create_preconditioner(A, settings(name, bs))
        p.compute()
        # every one of them is the exact inverse of a single diagonal block
        assert relerr(p.apply(r), np.linalg.solve(blk, r)) < 1e-12


def test_missing_diagonal_is_refused_by_preconditioners_only():
    m = matgen.SRMatrix(3, 1, np.array([0, 2, 3, 5], dtype=np.int32),
                        np.array([0, 1, 0, 1, 2], dtype=np.int32),
                        np.array([2.0, 1.0, 1.0, 1.0, 3.0]), None)
    m.diagind = matgen.find_diagind(m.browptr, m.bcolind, strict=False)
    assert m.diagind[1] == -1
    A = bb.SRMatrixView(m)
    x = np.array([1.0, 2.0, 3.0])
    assert relerr(A.apply(x), m.to_scipy() @ x) < 1e-14
    for name in ("jacobi", "sgs", "ilu0"):
        p = bb.SRFactory().create_preconditioner(A, settings(name, 1))
        with pytest.raises(RuntimeError, match="diagonal"):
            p.compute()


def test_dense_rows_beyond_tile_capacity():
    """One dense row and column (arrow matrix, 5000 entries in a row): the staged CSR kernels'
    tiles hold 3584 entries, so every scalar path has to take its long-row route."""
    import scipy.sparse as sp
    n = 5000
    rng = np.random.default_rng(3)
    a = sp.lil_matrix((n, n))
    a.setdiag(4.0 + rng.random(n))
    a[0, 1:] = rng.standard_normal(n - 1) * 1e-3
    a[1:, 0] = rng.standard_normal((n - 1, 1)) * 1e-3
    a[n - 1, 1:n - 1] = rng.standard_normal(n - 2) * 1e-3
    a[1:n - 1, n - 1] = rng.standard_normal((n - 2, 1)) * 1e-3
    m = matgen.from_scipy(a.tocsr())
    A = bb.SRMatrixView(m)
    x = rng.standard_normal(n)
    assert relerr(A.apply(x), orc().spmv(m, x)) < 1e-12
    r = rng.standard_normal(n)
    exact = orc().exact_ilu0(m)
    p = bb.SRFactory().create_preconditioner(A, settings("seqilu0", 1, nbuildsweeps=1, napplysweeps=1))
    p.compute()
    assert relerr(p.factor(), exact) < 1e-12
    pa = bb.SRFactory().create_preconditioner(A, settings("ilu0", 1, nbuildsweeps=20, napplysweeps=20))
    pa.compute()
    assert relerr(pa.factor(), exact) < 1e-10
    assert relerr(pa.apply(r), p.apply(r)) < 1e-10
    for name in ("sgs", "level_sgs"):
        q = bb.SRFactory().create_preconditioner(A, settings(name, 1, napplysweeps=20))
        q.compute()
        z = q.apply(r)
        assert np.isfinite(z).all()
