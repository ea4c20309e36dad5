"""GPU parity: Jacobi, asynchronous SGS, chaotic relaxation and level-scheduled SGS."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES, APPLY_INIT, LEVELS_CONTIGUOUS, LEVELS_DAG
from oracle import orc
from util import CASES, case, golden_outputs, relerr, SEED

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make(m, ptype, apply_init="init_jacobi", **kw):
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES[ptype], bs=m.bs,
                               blockstorage=1 if m.rowmajor else 0,
                               apply_inittype=APPLY_INIT[apply_init], **kw)
    return bb.SRFactory().create_preconditioner(m, s)


@pytest.mark.parametrize("key", CASES)
def test_jacobi(key):
    g, m = golden_outputs(), case(key)
    p = make(m, "jacobi")
    assert np.all(p.compute().f_info == 0)
    assert relerr(p.dblocks(), g[key + "_dblocks"]) < TOL
    assert relerr(p.apply(g[key + "_r"]), g[key + "_jacobi_apply"]) < TOL
    assert p.relaxationAvailable()
    # Jacobi relaxation is deterministic (synchronous): exact parity with the oracle
    d = orc().jacobi_setup(m)
    p.setApplyParams(maxits=4)
    x = p.apply_relax(g[key + "_r"], np.zeros(m.dim))
    assert relerr(x, orc().jacobi_relax(m, d, 4, g[key + "_r"], np.zeros(m.dim))) < TOL


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("mode", [LEVELS_DAG, LEVELS_CONTIGUOUS])
def test_level_sgs_apply_is_exact(key, mode):
    g, m = golden_outputs(), case(key)
    p = make(m, "level_sgs", level_mode=mode)
    p.compute()
    assert relerr(p.apply(g[key + "_r"]), g[key + "_sgs_apply"]) < TOL


@pytest.mark.parametrize("key", CASES)
def test_level_sgs_relax_contiguous_matches_sequential(key):
    """With the reference's contiguous levels the level-scheduled relaxation is the sequential one."""
    g, m = golden_outputs(), case(key)
    p = make(m, "level_sgs", level_mode=LEVELS_CONTIGUOUS)
    p.compute()
    p.setApplyParams(maxits=3)
    x = p.apply_relax(g[key + "_r"], np.zeros(m.dim))
    assert relerr(x, g[key + "_sgs_relax3"]) < 1e-11


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("ainit", ["init_zero", "init_jacobi"])
def test_async_sgs_converges_to_exact(key, ainit):
    g, m = golden_outputs(), case(key)
    p = make(m, "sgs", apply_init=ainit, napplysweeps=150)
    p.compute()
    assert relerr(p.apply(g[key + "_r"]), g[key + "_sgs_apply"]) < 1e-11


@pytest.mark.parametrize("mk", [lambda: matgen.poisson3d(12), lambda: matgen.block_stencil((16, 16), 4, SEED),
                                lambda: matgen.block_stencil((6, 6, 6), 5, SEED)])
def test_relaxations_converge_to_solution(mk):
    """Async SGS / GS relaxation and Jacobi relaxation drive A x = b to its solution on diagonally
    dominant matrices; same tolerance for the sequential oracle."""
    m = mk()
    rng = np.random.default_rng(SEED)
    xs = rng.standard_normal(m.dim)
    b = orc().spmv(m, xs)
    for ptype, its in (("sgs", 400), ("gs", 800)):
        p = make(m, ptype, napplysweeps=1)
        p.compute()
        p.setApplyParams(maxits=its)
        x = p.apply_relax(b, np.zeros(m.dim))
        assert relerr(x, xs) < 1e-8, ptype


def test_gs_apply_runs_sweeps_in_place():
    """ChaoticRelaxation::apply = napplysweeps forward sweeps starting from the contents of z."""
    m = case("2dcyl1_bsr4")
    p = make(m, "gs", napplysweeps=200)
    p.compute()
    rng = np.random.default_rng(2)
    xs = rng.standard_normal(m.dim)
    b = orc().spmv(m, xs)
    z = p.apply(b, np.zeros(m.dim))
    assert np.all(np.isfinite(z))


def test_noprec():
    m = case("2dcyl1_bsr4")
    p = make(m, "none")
    p.compute()
    r = np.arange(m.dim, dtype=np.float64)
    assert np.array_equal(p.apply(r), r)
    x = np.ones(m.dim)
    p.apply_relax(r, x)
    assert np.array_equal(x, np.ones(m.dim))
    assert not p.relaxationAvailable() and p.dim() == m.dim


def test_jacobi_relaxation_with_tolerance_checks():
    """ctol = true path of JacobiSRPreconditioner::apply_relax (solverops_jacobi.cpp:86-105): stops
    once the update is small relative to the first one."""
    m = matgen.block_stencil((12, 12), 4, SEED)
    p = make(m, "jacobi")
    p.compute()
    rng = np.random.default_rng(SEED)
    b = orc().spmv(m, rng.standard_normal(m.dim))
    d = orc().jacobi_setup(m)
    # emulate the reference loop on the host with the oracle's single relaxation step
    x = np.zeros(m.dim)
    ref0 = None
    for step in range(200):
        xn = orc().jacobi_relax(m, d, 1, b, x)
        diff = np.linalg.norm(xn - x)
        x = xn
        ref0 = diff if step == 0 else ref0
        if diff < 1e-50 or diff/ref0 < 1e-6 or diff/ref0 > 1e10:
            break
    p.setApplyParams(rtol=1e-6, atol=1e-50, dtol=1e10, ctol=True, maxits=200)
    xg = p.apply_relax(b, np.zeros(m.dim))
    assert relerr(xg, x) < 1e-11
    assert step < 199
