"""GPU parity of the asynchronous ILU(0) factorisation and application (scalar, bs=4 col/row major,
bs=5) through the C ABI.

Deterministic outputs (exact/"sequential" variants, initialisations, residuals) are held to 1e-12
relative; asynchronous outputs are checked the way the reference's own async tests do
(tests/solverops/CMakeLists.txt:6-67): convergence of the sweeps to the exact factorisation /
substitution, and stability of that fixed point."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES, FACT_INIT, APPLY_INIT
from oracle import orc
from util import CASES, case, golden_outputs, relerr, SEED

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make(m, ptype="ilu0", fact_init="init_original", apply_init="init_jacobi", **kw):
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES[ptype], bs=m.bs,
                               blockstorage=1 if m.rowmajor else 0,
                               fact_inittype=FACT_INIT[fact_init],
                               apply_inittype=APPLY_INIT[apply_init], **kw)
    return bb.SRFactory().create_preconditioner(m, s)


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("scaled", [False, True])
def test_exact_factor_and_apply_vs_golden(key, scaled):
    """BLASTED_SEQILU0: exact ILU(0) + exact triangular solves == the reference run sequentially."""
    g, m = golden_outputs(), case(key)
    tag = key + ("_scaled" if scaled else "")
    p = make(m, "seqilu0", scale=scaled, nbuildsweeps=1, napplysweeps=1)
    p.compute()
    assert relerr(p.factor(), g[tag + "_exact_ilu"]) < TOL
    assert relerr(p.apply(g[key + "_r"]), g[tag + "_ilu_apply"]) < TOL
    if scaled:
        assert relerr(p.scale_vector(), orc().scaling_vector(m)) < 1e-15


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("finit", ["init_original", "init_sgs", "init_zero"])
@pytest.mark.parametrize("scaled", [False, True])
def test_initialisations(key, finit, scaled):
    """Zero build sweeps leave the initial guess (diagonal blocks inverted afterwards for bs>1)."""
    m = case(key)
    if m.bs > 1 and finit == "init_zero":
        pytest.skip("zero blocks are singular (the reference produces inf/nan too)")
    O = orc()
    s = O.scaling_vector(m) if scaled else None
    want = O.ilu0_init(m, s, finit)
    O.ilu0_invert_diag(m, want)
    p = make(m, "ilu0", fact_init=finit, scale=scaled, nbuildsweeps=0)
    p.compute()
    assert relerr(p.factor(), want) < TOL


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("scaled", [False, True])
def test_precinfo_residuals(key, scaled):
    """PrecInfo: initial and final nonlinear residual sum|(A-LU)_S| and diagonal dominance."""
    O, m = orc(), case(key)
    s = O.scaling_vector(m) if scaled else None
    plist = O.ilu_positions(m)
    init = O.ilu0_init(m, s, "init_original")
    res0 = O.ilu0_nonlinear_res(m, plist, s, init)
    p = make(m, "seqilu0", scale=scaled, compute_precinfo=True)
    info = p.compute().f_info
    assert abs(info[1] - res0) <= 1e-12*res0
    exact = O.exact_ilu0(m, s)
    dd = O.diagonal_dominance(m, exact)
    assert np.allclose(info[[5, 4, 3, 2]], dd, rtol=1e-9, atol=1e-11)
    assert info[0] <= 1e-13*res0                   # remainder of the exact factorisation
    assert abs(p.ilu_residual() - info[0]) <= 1e-13*res0


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("finit", ["init_original", "init_sgs"])
def test_async_sweeps_converge_to_exact(key, finit):
    """Sweep convergence (tests/solverops/CMakeLists.txt:6-40: error < 1e-15 relative to the initial
    error within 150 sweeps) and the BASELINE criterion: nonlinear residual within 1e-10 (relative to
    ||A||) of the reference's after converged sweeps."""
    O, m = orc(), case(key)
    exact = O.exact_ilu0(m, None, invert_diag=True)
    plist = O.ilu_positions(m)
    ref_res = O.ilu0_nonlinear_res(m, plist, None, O.exact_ilu0(m)) / O.matrix_abs_sum(m)
    p = make(m, "ilu0", fact_init=finit, nbuildsweeps=150)
    p.compute()
    assert relerr(p.factor(), exact) < 1e-12
    assert p.ilu_residual()/O.matrix_abs_sum(m) <= ref_res + 1e-10


@pytest.mark.parametrize("key", ["2dcyl1_csr", "2dcyl1_bsr4", "synth_bsr5"])
def test_async_sweeps_monotone_and_fixed_point(key):
    """Error decreases with the sweep count, and extra sweeps at the fixed point change nothing
    (AsyncILU-ExactFixedPoint tests: 5 sweeps keep the change < 1e-16 relative)."""
    m = case(key)
    exact = orc().exact_ilu0(m, None, invert_diag=True)
    errs = []
    for nsw in (1, 3, 6, 12):
        p = make(m, "ilu0", nbuildsweeps=nsw)
        p.compute()
        errs.append(relerr(p.factor(), exact))
    assert errs[-1] < errs[0]
    pa, pb = make(m, "ilu0", nbuildsweeps=150), make(m, "ilu0", nbuildsweeps=155)
    pa.compute(); pb.compute()
    assert relerr(pa.factor(), pb.factor()) < 1e-15


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("ainit", ["init_zero", "init_jacobi"])
def test_async_apply_converges_to_exact_substitution(key, ainit):
    """Async triangular solves vs sequential substitution (AsyncILUTriangular tests)."""
    g, m = golden_outputs(), case(key)
    p = make(m, "sfilu0", apply_init=ainit, nbuildsweeps=1, napplysweeps=120)
    p.compute()
    assert relerr(p.apply(g[key + "_r"]), g[key + "_ilu_apply"]) < 1e-11


@pytest.mark.parametrize("key", ["2dcyl1_bsr4", "msc00726_csr"])
def test_apply_sweeps_bracketed_by_level_count(key):
    """The bracket of the asynchronous triangular sweeps: one sweep is what a synchronous
    (Jacobi-type) sweep gives at worst, and `nlevels` sweeps are the exact substitution whatever
    the interleaving - a row is final once the rows it reads are, so every sweep finalises at least
    one more dependency level (the iteration matrix is nilpotent of index nlevels).  Checked:
    k = nlevels sweeps reproduce the reference's sequential substitution to rounding, fewer sweeps
    give finite values, and on the diagonally dominant case the error shrinks with the sweeps."""
    g, m = golden_outputs(), case(key)
    r, want = g[key + "_r"], g[key + "_ilu_apply"]
    lv = make(m, "async_level_ilu0", nbuildsweeps=1)
    lv.compute()
    nlevels = len(lv.levels()[0]) - 1
    p = make(m, "sfilu0", napplysweeps=1)
    p.compute()
    errs = {}
    for k in (1, 2, 4, nlevels):
        p.set_sweeps(1, k)
        z = p.apply(r)
        assert np.all(np.isfinite(z))
        errs[k] = relerr(z, want)
    assert errs[nlevels] < 1e-12, errs
    if key.startswith("2dcyl1"):
        assert errs[4] < errs[1], errs


def test_level_scheduled_ilu_apply_both_modes():
    g = golden_outputs()
    for key in CASES:
        m = case(key)
        for mode in (0, 1):
            p = make(m, "async_level_ilu0", nbuildsweeps=150, level_mode=mode)
            p.compute()
            assert relerr(p.apply(g[key + "_r"]), g[key + "_ilu_apply"]) < 1e-12
    # the same substitution on the exact ("sequential") factor
    for key in CASES:
        p = make(case(key), "seqilu0")
        p.compute()
        assert relerr(p.apply(g[key + "_r"]), g[key + "_ilu_apply"]) < 1e-12


def test_reference_error_behaviour():
    m4, m1 = case("2dcyl1_bsr4"), case("2dcyl1_csr")
    p = make(m4, "ilu0")
    p.compute()
    assert not p.relaxationAvailable()
    with pytest.raises(RuntimeError, match="ILU relaxation not implemented!"):
        p.apply_relax(np.ones(m4.dim), np.zeros(m4.dim))
    # INIT_A_NONE makes the ILU apply drivers throw (solverops_ilu0.cpp:125-127,298-300)
    p = make(m1, "ilu0", apply_init="init_none")
    p.compute()
    with pytest.raises(RuntimeError, match="Invalid init type"):
        p.apply(np.ones(m1.dim))
    # factory argument errors (solverfactory.cpp:203-206,220-223)
    m3 = matgen.block_stencil((4, 4), 3, 1)
    with pytest.raises(ValueError, match="not supported for column major"):
        make(m3, "ilu0")
    m5r = matgen.block_stencil((4, 4), 5, 1, rowmajor=True)
    with pytest.raises(ValueError, match="not supported for row major"):
        make(m5r, "ilu0")
    with pytest.raises(ValueError, match="Invalid preconditioner"):
        bb.SRFactory().create_preconditioner(m1, bb.AsyncSolverSettings(prectype=7, bs=1))
    with pytest.raises(RuntimeError, match="before compute"):
        make(m1, "ilu0").apply(np.ones(m1.dim))
    assert p.dim() == m1.dim


def test_recompute_with_new_values_same_pattern():
    """compute() again after the values changed, pattern cached (solverops_ilu0.hpp:53-56)."""
    m = case("2dcyl1_bsr4")
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["seqilu0"], bs=4)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    f1 = p.factor()
    m2 = matgen.SRMatrix(m.nbrows, 4, m.browptr, m.bcolind, m.vals*1.5, m.diagind)
    view.update_values(m2.vals)
    p.compute()
    want = orc().exact_ilu0(m2, None, invert_diag=True)
    assert relerr(p.factor(), want) < TOL
    assert relerr(p.factor(), f1) > 1e-3


@pytest.mark.parametrize("mk", [lambda: matgen.poisson3d(16), lambda: matgen.poisson3d(10, 27),
                                lambda: matgen.block_stencil((24, 20), 4, SEED),
                                lambda: matgen.block_stencil((8, 7, 6), 5, SEED)])
def test_synthetic_exact_and_async(mk):
    m = mk()
    O = orc()
    scaled = m.bs == 1 and m.avg_row_len > 8 if hasattr(m, "avg_row_len") else False
    exact = O.exact_ilu0(m, None, invert_diag=True)
    p = make(m, "seqilu0")
    p.compute()
    assert relerr(p.factor(), exact) < TOL
    r = np.random.default_rng(SEED).standard_normal(m.dim)
    assert relerr(p.apply(r), O.ilu0_apply(m, exact, None, 1, "init_zero", r)) < TOL


def test_one_launch_exact_factorisation_small_cases():
    """The one-launch exact scalar factorisation is only selected from a million rows up; force it
    on the fixtures (fresh process: the switch is read once) and compare with the oracle."""
    import os, subprocess, sys
    code = r'''
import sys, os
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import blasted_b200 as bb
from blasted_b200.solverfactory import SOLVER_TYPES
from oracle import orc
from util import case, relerr
for key, scale in (("2dcyl1_csr", False), ("2dcyl1_csr", True), ("msc00726_csr", True)):
    m = case(key)
    sv = orc().scaling_vector(m) if scale else None
    exact = orc().exact_ilu0(m, sv)
    p = bb.SRFactory().create_preconditioner(bb.SRMatrixView(m), bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["sfilu0"], bs=1, scale=scale, nbuildsweeps=1))
    p.compute()
    assert relerr(p.factor(), exact) < 1e-12, key
    r = np.cos(np.arange(m.dim))
    q = bb.SRFactory().create_preconditioner(bb.SRMatrixView(m), bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["seqilu0"], bs=1, scale=scale, nbuildsweeps=1))
    q.compute()
    assert np.array_equal(q.factor(), p.factor())
    assert np.isfinite(q.apply(r)).all()
print("ok")
'''
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    env = dict(os.environ, B200_EXACT_ONE_LAUNCH="1", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("bs", [4, 5])
@pytest.mark.parametrize("scaled", [False, True])
def test_block_factor_on_a_general_pattern(bs, scaled):
    """A block pattern in which lower AND strict upper entries have products (a 27-point operator
    re-blocked): the staged bs = 5 launches do not apply (they need the face-neighbour structure),
    the generic launches, the one-launch exact factorisation and the sweeps must all agree with
    the sequential pass of the oracle."""
    m = matgen.csr_to_bsr(matgen.poisson3d(0, 27, dims=(2*bs, 5, 4)), bs)
    rng = np.random.default_rng(SEED + bs)
    m.vals = m.vals + 0.05*rng.standard_normal(m.vals.shape)          # unsymmetric values, same pattern
    O = orc()
    s = O.scaling_vector(m) if scaled else None
    want = O.exact_ilu0(m, s, invert_diag=True)
    stats = None
    for ptype, kw in (("sfilu0", dict(nbuildsweeps=1)), ("ilu0", dict(nbuildsweeps=60))):
        p = make(m, ptype, scale=scaled, **kw)
        p.compute()
        assert relerr(p.factor(), want) < 1e-11, ptype
        stats = p.pattern_stats()
    assert stats["npos_l"] > 0 and stats["nuwork"] > m.nbrows        # products everywhere
    # converged asynchronous triangular sweeps on that factor == the exact substitution
    r = rng.standard_normal(m.dim)
    q = make(m, "seqilu0", scale=scaled)
    q.compute()
    a = make(m, "sfilu0", scale=scaled, nbuildsweeps=1, napplysweeps=80)
    a.compute()
    assert relerr(a.apply(r), q.apply(r)) < 1e-11


@pytest.mark.parametrize("key", ["2dcyl1_bsr4", "2dcyl1_bsr4r", "synth_bsr5", "2dcyl1_csr"])
def test_compute_with_new_host_values_equals_update_then_compute(key):
    """b200_prec_compute_host (values uploaded chunk by chunk, layout conversion and initial guess
    behind the copies) gives bit for bit what update_values + compute gives - on exact types, so
    that the comparison is deterministic."""
    m = case(key)
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["sfilu0"], bs=m.bs, blockstorage=1 if m.rowmajor else 0)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    rng = np.random.default_rng(SEED + 11)
    newvals = m.vals*(1.0 + 0.01*rng.standard_normal(m.vals.shape))
    p.compute(newvals)
    got = p.factor()
    view2 = bb.SRMatrixView(m)
    q = bb.SRFactory().create_preconditioner(view2, s)
    q.compute()
    view2.update_values(newvals)
    q.compute()
    assert np.array_equal(got, q.factor())
    # and the asynchronous type after the pipelined upload converges to the same factor
    a = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["ilu0"], bs=m.bs, blockstorage=1 if m.rowmajor else 0, nbuildsweeps=60))
    a.compute()
    a.compute(newvals)
    assert relerr(a.factor(), got) < 1e-11
