"""CPU tests: pin the plain-C oracle (oracle/blasted_oracle.c) against the golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py), and - when oracle/_ref is present - against
the reference library itself, live."""
import numpy as np
import pytest

from oracle import orc, ref, have_ref
from util import CASES, case, golden_outputs, golden_matrices, fixture_csr, relerr

TOL = 1e-12       # fp64 deterministic paths (BASELINE.json north_star)


@pytest.mark.parametrize("key", CASES)
def test_spmv_gemv3_vs_golden(key):
    O, g, m = orc(), golden_outputs(), case(key)
    r = g[key + "_r"]
    assert relerr(O.spmv(m, r), g[key + "_spmv"]) < TOL
    y = np.cos(np.arange(m.dim))
    assert relerr(O.gemv3(m, 0.3, r, -1.2, y), g[key + "_gemv3"]) < TOL


@pytest.mark.parametrize("name,bs", [("DK01R", 1), ("small_block3", 1), ("2dcyl1", 1), ("msc00726", 1)])
def test_spmv_fixture_products(name, bs):
    """b = A x for the reference's stored x/b pairs (tests/mat_ops/testcsrmatrix.cpp:34-36)."""
    gm = golden_matrices()
    m = fixture_csr(name, strict=False)
    b = orc().spmv(m, gm[name + "_x"])
    assert relerr(b, gm[name + "_b"]) < 1e-13


@pytest.mark.parametrize("key", CASES)
def test_ilu_positions_bit_exact(key):
    g, m = golden_outputs(), case(key)
    posptr, lowerp, upperp = orc().ilu_positions(m)
    assert np.array_equal(posptr, g[key + "_posptr"])
    assert np.array_equal(lowerp, g[key + "_lowerp"])
    assert np.array_equal(upperp, g[key + "_upperp"])


@pytest.mark.parametrize("key", CASES)
def test_levels_bit_exact(key):
    g, m = golden_outputs(), case(key)
    lv = orc().compute_levels(m)
    assert np.array_equal(lv, g[key + "_levels"])
    # the reference's own property test (tests/mat_ops/testlevelschedule.cpp:25-37)
    for l in range(len(lv) - 1):
        for i in range(lv[l], lv[l+1]):
            cols = m.bcolind[m.browptr[i]:m.browptr[i+1]]
            assert not np.any((cols >= lv[l]) & (cols < i))


def test_levels_nonsymmetric_raises():
    from blasted_b200 import matgen
    import scipy.sparse as sp
    a = sp.csr_matrix(np.array([[2., 1, 0], [0, 2, 0], [0, 1, 2]]))
    with pytest.raises(RuntimeError):
        orc().compute_levels(matgen.from_scipy(a))


@pytest.mark.parametrize("key", CASES)
@pytest.mark.parametrize("scaled", [False, True])
def test_exact_ilu_and_apply_vs_golden(key, scaled):
    O, g, m = orc(), golden_outputs(), case(key)
    tag = key + ("_scaled" if scaled else "")
    s = O.scaling_vector(m) if scaled else None
    ilu = O.exact_ilu0(m, s, invert_diag=True)
    assert relerr(ilu, g[tag + "_exact_ilu"]) < TOL
    z = O.ilu0_apply(m, ilu, s, 1, "init_jacobi", g[key + "_r"])
    assert relerr(z, g[tag + "_ilu_apply"]) < TOL


@pytest.mark.parametrize("key", CASES)
def test_precinfo_vs_golden(key):
    O, g, m = orc(), golden_outputs(), case(key)
    plist = O.ilu_positions(m)
    ilu = O.ilu0_init(m, None, "init_original")
    info = g[key + "_precinfo"]
    assert abs(O.ilu0_nonlinear_res(m, plist, None, ilu) - info[1]) <= 1e-12*info[1]
    O.ilu0_sweeps(m, plist, None, 1, ilu)
    # remainder of the exact factorisation: rounding-level relative to the initial remainder
    assert O.ilu0_nonlinear_res(m, plist, None, ilu) < 1e-13*info[1]
    dd = O.diagonal_dominance(m, ilu)
    assert np.allclose(dd, info[[5, 4, 3, 2]], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("key", CASES)
def test_sgs_jacobi_vs_golden(key):
    O, g, m = orc(), golden_outputs(), case(key)
    r = g[key + "_r"]
    d = O.jacobi_setup(m)
    assert relerr(d, g[key + "_dblocks"]) < TOL
    assert relerr(O.jacobi_apply(m, d, r), g[key + "_jacobi_apply"]) < TOL
    assert relerr(O.sgs_apply(m, d, 1, "init_jacobi", r), g[key + "_sgs_apply"]) < TOL
    assert relerr(O.sgs_relax(m, d, 3, r, np.zeros(m.dim)), g[key + "_sgs_relax3"]) < TOL


def test_sequential_sweep_is_fixed_point():
    """Exact-ILU fixed point: further sequential sweeps do not change the factor
    (tests/solverops/CMakeLists.txt:51-67, bound 1e-16 relative)."""
    O, m = orc(), case("2dcyl1_bsr4")
    plist = O.ilu_positions(m)
    ilu = O.exact_ilu0(m)
    again = O.ilu0_sweeps(m, plist, None, 5, ilu.copy())
    assert relerr(again, ilu) < 1e-14


def test_synchronous_sweeps_converge_to_exact():
    """Fully synchronous (Jacobi-type) sweeps - the other extreme of chaotic iteration - reach the
    same fixed point (7-point Poisson converges unscaled)."""
    from blasted_b200 import matgen
    O, m = orc(), matgen.poisson3d(8)
    plist = O.ilu_positions(m)
    exact = O.exact_ilu0(m)
    ilu = O.ilu0_init(m, None, "init_original")
    for _ in range(40):
        ilu = O.ilu0_sweep_synchronous(m, plist, None, ilu)
    assert relerr(ilu, exact) < 1e-12


def test_block_inverse():
    rng = np.random.default_rng(3)
    for bs in (4, 5):
        a = rng.standard_normal((bs, bs)) + 3*np.eye(bs)
        inv = orc().block_inverse(bs, True, a.ravel()).reshape(bs, bs)
        assert np.allclose(inv @ a, np.eye(bs), atol=1e-13)
        invc = orc().block_inverse(bs, False, a.T.ravel()).reshape(bs, bs).T
        assert np.allclose(invc, inv, atol=1e-13)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (reference tree absent)")
@pytest.mark.parametrize("key", ["2dcyl1_bsr4", "synth_bsr5", "msc00726_csr"])
def test_oracle_vs_live_reference(key):
    """Same checks against the reference library itself, including init_sgs and threaded runs."""
    O, R, m = orc(), ref(), case(key)
    R.set_num_threads(1)
    rng = np.random.default_rng(7)
    r = rng.standard_normal(m.dim)
    for scale in (False, True):
        if m.bs == 1 and scale:
            continue            # scalar scaled init_sgs reads out of bounds in the reference
        p = R.prec(m, "seqilu0", scale=scale, nbuildsweeps=2, napplysweeps=2, fact_init="init_sgs",
                   apply_init="init_zero")
        p.compute()
        s = O.scaling_vector(m) if scale else None
        plist = O.ilu_positions(m)
        ilu = O.ilu0_init(m, s, "init_sgs")
        O.ilu0_sweeps(m, plist, s, 2, ilu)
        O.ilu0_invert_diag(m, ilu)
        assert relerr(ilu, p.factor()) < TOL
        assert relerr(O.ilu0_apply(m, ilu, s, 2, "init_zero", r), p.apply(r)) < TOL
        p.close()
