"""Device counterpart of the reference's stand-alone test driver `runsolvetest`
(tests/runsolvetest.cpp:25, tests/testsolve.cpp:29-187): read a Matrix-Market system, build the
requested preconditioner through the factory, solve with BiCGSTAB / GCR / Richardson (plus FGMRES
proper) on the GPU and check the solution against a reference vector.

Same option names as the reference's boost::program_options set (tests/testsolve.cpp:133-187):

    python -m blasted_b200.testsolve --solver_type bcgs --preconditioner_type ilu0 \
        --mat_type bsr --block_size 4 --storage_order colmajor --build_sweeps 10 --apply_sweeps 15 \
        --mat_file 2dcyl1.mtx --b_file 2dcyl1_b.mtx --x_file 2dcyl1_x.mtx --test_tol 1e-4

Matrix-Market ingestion (COOMatrix::readMatrixMarket + getSRMatrixFromCOO, src/coomatrix.cpp:189,
:427) is host I/O and done with scipy; everything after it runs through the C ABI.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np


def read_matrix_market(path: str, bs: int = 1, rowmajor: bool = False):
    """COOMatrix::readMatrixMarket + getSRMatrixFromCOO<bs> (src/coomatrix.cpp:189-462)."""
    import scipy.io as sio
    from . import matgen
    return matgen.from_scipy(sio.mmread(path), bs, rowmajor)


def read_dense_matrix_market(path: str) -> np.ndarray:
    """readDenseMatrixMarket (include/coomatrix.hpp)."""
    import scipy.io as sio
    return np.asarray(sio.mmread(path), dtype=np.float64).ravel()


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(description="Options available for test solve with device solvers")
    ap.add_argument("--solver_type", default="bcgs", help="[bcgs/richardson/gcr/fgmres]")
    ap.add_argument("--preconditioner_type", default="jacobi",
                    help="[none/jacobi/gs/sgs/level_sgs/ilu0/seqilu0/sfilu0/sapilu0/async_level_ilu0]")
    ap.add_argument("--fact_init_type", default="init_original")
    ap.add_argument("--apply_init_type", default="init_zero")
    ap.add_argument("--mat_type", default="csr", help="[csr/bsr]")
    ap.add_argument("--block_size", type=int, default=4)
    ap.add_argument("--storage_order", default="colmajor", help="[rowmajor/colmajor]")
    ap.add_argument("--test_tol", type=float, default=1e-4)
    ap.add_argument("--solver_tol", type=float, default=1e-6)
    ap.add_argument("--max_iter", type=int, default=1000)
    ap.add_argument("--solver_restart", type=int, default=30)
    ap.add_argument("--build_sweeps", type=int, default=1)
    ap.add_argument("--apply_sweeps", type=int, default=1)
    ap.add_argument("--thread_chunk_size", type=int, default=256)
    ap.add_argument("--use_symmetric_scaling", action="store_true",
                    help="-blasted_use_symmetric_scaling of the PETSc interface (doc/user-doc.md:6-28)")
    ap.add_argument("--mat_file", required=True)
    ap.add_argument("--b_file", required=True)
    ap.add_argument("--x_file", default="NONE")
    return ap


def test_solve(params) -> int:
    import blasted_b200 as bb
    from .solverfactory import getFactInitFromString, getApplyInitFromString, ROWMAJOR, COLMAJOR

    bs = params.block_size if params.mat_type == "bsr" else 1
    rowmajor = params.storage_order == "rowmajor"
    print(f"Inputs: Solver = {params.solver_type}, Prec = {params.preconditioner_type}, order = "
          f"{params.storage_order}, test tol = {params.test_tol}, tolerance = {params.solver_tol} "
          f"maxiter = {params.max_iter},\n  Num build sweeps = {params.build_sweeps}, "
          f"num apply sweeps = {params.apply_sweeps}")
    m = read_matrix_market(params.mat_file, bs, rowmajor)
    b = read_dense_matrix_market(params.b_file)
    print(f"Read matrix with {m.nbrows} (block-)rows, and {m.nnzb} nonzero blocks, with block size {bs}")
    print(f"Read RHS vector with {len(b)} rows")

    mat = bb.SRMatrixView(m)
    fctry = bb.SRFactory()
    aparams = bb.AsyncSolverSettings(
        prectype=fctry.solverTypeFromString(params.preconditioner_type), bs=bs,
        blockstorage=ROWMAJOR if rowmajor else COLMAJOR, relax=False,
        thread_chunk_size=params.thread_chunk_size, scale=params.use_symmetric_scaling,
        nbuildsweeps=params.build_sweeps, napplysweeps=params.apply_sweeps,
        fact_inittype=getFactInitFromString(params.fact_init_type),
        apply_inittype=getApplyInitFromString(params.apply_init_type))
    prec = fctry.create_preconditioner(mat, aparams)
    prec.compute()

    if params.solver_type == "richardson":
        solver = bb.RichardsonSolver(mat, prec)
    elif params.solver_type == "bcgs":
        solver = bb.BiCGSTAB(mat, prec)
    elif params.solver_type == "gcr":
        solver = bb.GCR(mat, prec, params.solver_restart)
    elif params.solver_type == "fgmres":
        solver = bb.FGMRES(mat, prec, params.solver_restart)
    else:
        print(" ! Invalid solver option!")
        return 2
    solver.setParams(params.solver_tol, params.max_iter)
    print("Starting solve ")
    x = np.zeros(mat.dim())
    info = solver.solve(b, x)
    print(f"  Final rel res norm = {info.resnorm/info.bnorm:g}")
    print(f" Num iters = {info.iters}")
    if params.x_file != "NONE":
        ans = read_dense_matrix_market(params.x_file)
        l2norm = float(np.sqrt(np.sum((x - ans)**2)))
        print(f" L2 norm of error = {l2norm:g}")
        if not l2norm < params.test_tol:
            print(" ! error above test_tol")
            return 1
    return 0


def main(argv=None) -> int:
    return test_solve(build_parser().parse_args(argv))


if __name__ == "__main__":
    sys.exit(main())
