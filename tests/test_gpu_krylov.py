"""GPU: Krylov drivers + preconditioners vs the reference's iteration counts (golden, sequential
reference runs of tests/solvers.cpp; BASELINE.json: iteration counts within 5 %)."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200.solverfactory import SOLVER_TYPES
from oracle import orc
from util import case, golden_outputs, golden_matrices, relerr

pytestmark = pytest.mark.gpu

# iteration counts of the unmodified reference over OpenMP thread counts 1..16 where they vary
# through the reduction order of the dot products alone (measured on the GPU box's host by
# tools/its_probe.py with oracle/_ref; min, max)
REF_THREAD_SPREAD = {("msc00726_csr", "jacobi", "bicgstab"): (71, 76)}

# device type whose result equals the reference's sequential preconditioner
EXACT = {"seqilu0": "seqilu0", "sgs": "level_sgs", "jacobi": "jacobi"}


def solve(m, ptype, solver, b, tol=1e-10, **kw):
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES[ptype], bs=m.bs,
                               blockstorage=1 if m.rowmajor else 0, **kw)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    sol = bb.BiCGSTAB(view, p) if solver == "bicgstab" else bb.GCR(view, p, 30)
    sol.setParams(tol, 2000)
    x = np.zeros(m.dim)
    info = sol.solve(b, x)
    return x, info


@pytest.mark.parametrize("key", ["2dcyl1_csr", "2dcyl1_bsr4", "2dcyl1_bsr4r", "msc00726_csr"])
@pytest.mark.parametrize("prec", ["seqilu0", "sgs", "jacobi"])
@pytest.mark.parametrize("solver", ["bicgstab", "gcr"])
def test_iteration_counts_match_sequential_reference(key, prec, solver):
    g, gm, m = golden_outputs(), golden_matrices(), case(key)
    b = gm[key.split("_")[0] + "_b"]
    want_its = int(g[f"its_{key}_{prec}_{solver}"][0])
    x, info = solve(m, EXACT[prec], solver, b)
    assert info.converged or info.resnorm/info.bnorm < 1e-10
    # within 5 % (+1 iteration of slack for tiny counts): rounding differs, the algorithm does not.
    # One case needs the reference's OWN spread as the yardstick: msc00726 (cond ~4e5) with the
    # weakest preconditioner makes BiCGSTAB's count depend on the rounding of the dot products, and
    # the reference's OpenMP reductions reorder them with the thread count - the unmodified
    # reference needs 71, 76, 72, 76, 76, 76, 74, 76 iterations with 1, 2, 3, 4, 6, 8, 12, 16 threads
    # (Jacobi itself does not depend on threads; tools/its_probe.py, profiles/its_probe_r02.log).
    # The device count has to lie within 5 % of that range; every other case within 5 % of the
    # sequential count.
    lo = hi = want_its
    if (key, prec, solver) == ("msc00726_csr", "jacobi", "bicgstab"):
        lo, hi = REF_THREAD_SPREAD[(key, prec, solver)]
        assert lo <= want_its <= hi
    slack = lambda w: max(1, int(np.ceil(0.05*w)))
    assert lo - slack(lo) <= info.iters <= hi + slack(hi), (info.iters, want_its, lo, hi)
    # and the solution solves the system
    res = np.linalg.norm(b - orc().spmv(m, x))/np.linalg.norm(b)
    assert res < 5e-10


@pytest.mark.parametrize("key,nb,na,scale", [("2dcyl1_bsr4", 10, 30, False), ("2dcyl1_csr", 30, 60, False),
                                             ("2dcyl1_bsr4", 10, 30, True), ("msc00726_csr", 30, 60, True)])
def test_async_ilu0_iterations_within_5_percent(key, nb, na, scale):
    """Async ILU(0) with converged sweeps reproduces the sequential iteration count
    (reference threaded test ThreadedBSR4ILU0Colmajor uses sweeps 10/15, tests/CMakeLists.txt:166-173)."""
    g, gm, m = golden_outputs(), golden_matrices(), case(key)
    b = gm[key.split("_")[0] + "_b"]
    want = int(g[f"its_{key}_seqilu0_bicgstab"][0])
    # msc00726 is far from diagonally dominant: on ~10^5 concurrent threads the chaotic ILU iteration
    # is close to a synchronous (Jacobi-type) one, which overflows for this matrix unless the
    # reference's own symmetric scaling option is switched on (measured in
    # tools/conv_study.py: unscaled needs ~150 sweeps, scaled converges in < 30)
    x, info = solve(m, "ilu0", "bicgstab", b, nbuildsweeps=nb, napplysweeps=na, scale=scale)
    assert abs(info.iters - want) <= max(1, int(np.ceil(0.05*want))), (info.iters, want)


def test_richardson_and_noprec():
    gm, m = golden_matrices(), case("2dcyl1_bsr4")
    b = gm["2dcyl1_b"]
    view = bb.SRMatrixView(m)
    p = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["level_sgs"], bs=4))
    p.compute()
    sol = bb.RichardsonSolver(view, p)
    sol.setParams(1e-6, 500)
    x = np.zeros(m.dim)
    info = sol.solve(b, x)
    assert info.iters > 0 and np.isfinite(info.resnorm)


@pytest.mark.parametrize("key", ["2dcyl1_bsr4", "2dcyl1_csr", "msc00726_csr"])
@pytest.mark.parametrize("prec", ["seqilu0", "level_sgs"])
def test_fgmres_matches_gcr(key, prec):
    """FGMRES(30) proper and the reference's GCR(30) are the same method in exact arithmetic
    (tests/solvers.hpp:108-110): same iteration count (within 5 %), both solve the system."""
    gm, m = golden_matrices(), case(key)
    b = gm[key.split("_")[0] + "_b"]
    view = bb.SRMatrixView(m)
    p = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES[prec], bs=m.bs, blockstorage=1 if m.rowmajor else 0))
    p.compute()
    its = {}
    for name, cls in (("gcr", bb.GCR), ("fgmres", bb.FGMRES)):
        sol = cls(view, p, 30)
        sol.setParams(1e-10, 2000)
        x = np.zeros(m.dim)
        info = sol.solve(b, x)
        its[name] = info.iters
        assert np.linalg.norm(b - orc().spmv(m, x))/np.linalg.norm(b) < 5e-10
    assert abs(its["fgmres"] - its["gcr"]) <= max(1, int(np.ceil(0.05*its["gcr"]))), its


def test_runsolvetest_driver(tmp_path):
    """The device counterpart of `runsolvetest` with the options of the reference's CTest entries
    (tests/CMakeLists.txt:86-173, e.g. BSR4ILU0Colmajor / ThreadedBSR4ILU0Colmajor sweeps 10/15)."""
    import scipy.io as sio
    import scipy.sparse as sp
    from blasted_b200 import testsolve
    gm = golden_matrices()
    m = case("2dcyl1_csr")
    a = sp.csr_matrix((m.vals, m.bcolind, m.browptr), shape=(m.nbrows, m.nbrows))
    sio.mmwrite(str(tmp_path/"a.mtx"), a, precision=17)
    sio.mmwrite(str(tmp_path/"b.mtx"), gm["2dcyl1_b"].reshape(-1, 1), precision=17)
    sio.mmwrite(str(tmp_path/"x.mtx"), gm["2dcyl1_x"].reshape(-1, 1), precision=17)
    base = ["--mat_file", str(tmp_path/"a.mtx"), "--b_file", str(tmp_path/"b.mtx"),
            "--x_file", str(tmp_path/"x.mtx"), "--solver_tol", "1e-10", "--test_tol", "1e-6"]
    assert testsolve.main(base + ["--solver_type", "bcgs", "--preconditioner_type", "ilu0", "--mat_type",
                                  "bsr", "--block_size", "4", "--build_sweeps", "10",
                                  "--apply_sweeps", "15", "--apply_init_type", "init_jacobi"]) == 0
    assert testsolve.main(base + ["--solver_type", "gcr", "--preconditioner_type", "level_sgs",
                                  "--mat_type", "csr"]) == 0
    assert testsolve.main(base + ["--solver_type", "fgmres", "--preconditioner_type", "seqilu0",
                                  "--mat_type", "bsr", "--storage_order", "rowmajor"]) == 0


@pytest.mark.parametrize("solver", ["gcr", "fgmres"])
def test_in_place_preconditioner_starts_from_zero_in_every_solve(solver):
    """The chaotic GS relaxation used as a preconditioner sweeps in place on whatever its output
    vector holds on entry (relaxation_chaotic.cpp:22-45).  The restarted drivers keep their work
    vectors between solves; like the reference's value-initialised vectors (tests/solvers.cpp:264-270)
    they must start from zero in every solve: two consecutive solves give the same iteration count
    and the same (finite) solution."""
    import torch
    m = case("2dcyl1_bsr4")
    view = bb.SRMatrixView(m)
    p = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["gs"], bs=4, napplysweeps=3))
    p.compute()
    b = torch.from_numpy(golden_matrices()["2dcyl1_b"]).cuda()
    runs = []
    for rep in range(2):
        sol = bb.GCR(view, p, 30) if solver == "gcr" else bb.FGMRES(view, p, 30)
        sol.setParams(1e-8, 2000)
        x = torch.zeros_like(b)
        info = sol.solve(b, x)
        assert info.converged and bool(torch.isfinite(x).all())
        runs.append((info.iters, x.cpu().numpy()))
    # asynchronous sweeps: counts agree to a few iterations, solutions to the solve tolerance
    assert abs(runs[0][0] - runs[1][0]) <= max(2, runs[0][0]//10), (runs[0][0], runs[1][0])
    assert relerr(runs[0][1], runs[1][1]) < 1e-5


def test_solve_rejects_mismatched_streams_and_reports_deferred_errors():
    """b200_solve refuses a preconditioner and a matrix on different streams (their launches would
    not be ordered); b200_prec_check reports nothing after healthy applies."""
    import ctypes as C
    import torch
    from blasted_b200._lib import lib
    m = case("2dcyl1_csr")
    view = bb.SRMatrixView(m)
    p = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES["seqilu0"], bs=1))
    p.compute()
    r = torch.from_numpy(golden_outputs()["2dcyl1_csr_r"]).cuda()
    z = torch.empty_like(r)
    p.apply(r, z)
    assert lib.b200_prec_check(p._h) == 0
    st = torch.cuda.Stream()
    lib.b200_prec_set_stream(p._h, C.c_void_p(st.cuda_stream))
    sol = bb.GCR(view, p, 30)
    sol.setParams(1e-8, 100)
    with pytest.raises(RuntimeError, match="different streams"):
        sol.solve(r, torch.zeros_like(r))
    lib.b200_prec_set_stream(p._h, C.c_void_p(0))
    info = sol.solve(r, torch.zeros_like(r))
    assert info.converged
