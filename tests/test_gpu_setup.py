"""GPU parity of the setup built on the device: ILU position lists and level schedules are
integer outputs and must be BIT-EXACT (BASELINE.json north_star)."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES, LEVELS_CONTIGUOUS, LEVELS_DAG
from oracle import orc
from util import CASES, case, golden_outputs, SEED

pytestmark = pytest.mark.gpu


def make(m, ptype, **kw):
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES[ptype], bs=m.bs,
                               blockstorage=1 if m.rowmajor else 0, **kw)
    return bb.SRFactory().create_preconditioner(m, s)


@pytest.mark.parametrize("key", CASES)
def test_ilu_positions_fixtures(key):
    g, m = golden_outputs(), case(key)
    p = make(m, "ilu0")
    p.compute()
    posptr, lowerp, upperp = p.ilu_positions()
    assert np.array_equal(posptr, g[key + "_posptr"])
    assert np.array_equal(lowerp, g[key + "_lowerp"])
    assert np.array_equal(upperp, g[key + "_upperp"])


@pytest.mark.parametrize("mk", [lambda: matgen.poisson3d(24), lambda: matgen.poisson3d(14, 27),
                                lambda: matgen.block_stencil((40, 30), 4, SEED),
                                lambda: matgen.poisson3d(1), lambda: matgen.poisson2d(5, 1)])
def test_ilu_positions_synthetic(mk):
    m = mk()
    p = make(m, "ilu0")
    p.compute()
    got, want = p.ilu_positions(), orc().ilu_positions(m)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("key", CASES)
def test_contiguous_levels_fixtures(key):
    """computeLevels output, and the reference's property test
    (tests/mat_ops/testlevelschedule.cpp:25-37; 2dcyl1/bs4 -> 43 levels)."""
    g, m = golden_outputs(), case(key)
    p = make(m, "level_sgs", level_mode=LEVELS_CONTIGUOUS)
    p.compute()
    ptr, rows = p.levels()
    assert np.array_equal(ptr, g[key + "_levels"])
    assert np.array_equal(rows, np.arange(m.nbrows))
    if key == "2dcyl1_bsr4":
        assert len(ptr) - 1 == 43


@pytest.mark.parametrize("mk", [lambda: matgen.poisson3d(14), lambda: matgen.poisson3d(14, 27),
                                lambda: matgen.poisson3d(21, dims=(40, 30, 9))])
def test_contiguous_levels_synthetic(mk):
    m = mk()
    p = make(m, "async_level_ilu0", level_mode=LEVELS_CONTIGUOUS)
    p.compute()
    ptr, _ = p.levels()
    assert np.array_equal(ptr, orc().compute_levels(m))


@pytest.mark.parametrize("mk,expect", [(lambda: matgen.poisson3d(10), 3*10 - 2),
                                       (lambda: matgen.poisson3d(8, 27), 7*8 - 6),
                                       (lambda: case("2dcyl1_bsr4"), None),
                                       (lambda: case("msc00726_csr"), None)])
def test_dag_levels(mk, expect):
    m = mk()
    p = make(m, "level_sgs", level_mode=LEVELS_DAG)
    p.compute()
    ptr, rows = p.levels()
    nlev, lv = orc().dag_levels(m)
    assert len(ptr) - 1 == nlev
    if expect is not None:
        assert nlev == expect        # 7-pt 3n-2, 27-pt 7n-6 wavefronts
    order = np.argsort(lv, kind="stable")
    assert np.array_equal(rows, order.astype(np.int32))
    assert np.array_equal(ptr, np.concatenate([[0], np.cumsum(np.bincount(lv))]).astype(np.int32))


@pytest.mark.parametrize("n,density,seed", [(1, 0.0, 0), (7, 0.0, 1), (33, 0.2, 2), (500, 0.01, 3),
                                            (5000, 0.0008, 4), (20000, 0.0001, 5)])
def test_contiguous_levels_random_symmetric_patterns(n, density, seed):
    """Scatter-min + suffix-min scan + pointer doubling against the row-by-row walk of computeLevels
    (src/levelschedule.cpp:12-71) on irregular structurally symmetric patterns (diagonal-only and
    1x1 included: one level)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(SEED + seed)
    a = sp.random(n, n, density=density, random_state=rng, format="csr")
    a = (a + a.T + sp.identity(n) * (n + 1.0)).tocsr()
    a.sort_indices()
    m = matgen.from_scipy(a)
    p = make(m, "level_sgs", level_mode=LEVELS_CONTIGUOUS)
    p.compute()
    ptr, rows = p.levels()
    assert np.array_equal(ptr, orc().compute_levels(m))
    assert np.array_equal(rows, np.arange(n))


def test_nonsymmetric_pattern_rejected_for_contiguous_levels():
    import scipy.sparse as sp
    a = sp.csr_matrix(np.array([[2., 1, 0], [0, 2, 0], [0, 1, 2]]))
    p = make(matgen.from_scipy(a), "level_sgs", level_mode=LEVELS_CONTIGUOUS)
    with pytest.raises(RuntimeError, match="Faulty dependency list"):
        p.compute()
