"""Device vs golden (sequential reference) Krylov iteration counts on the fixtures, and the
reference's own spread over OpenMP thread counts where its preconditioner does not depend on them
(development tool behind the tolerance of tests/test_gpu_krylov.py)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np
import blasted_b200 as bb
from blasted_b200.solverfactory import SOLVER_TYPES
from util import case, golden_outputs, golden_matrices
from oracle import have_ref, ref

EXACT = {"seqilu0": "seqilu0", "sgs": "level_sgs", "jacobi": "jacobi"}
g, gm = golden_outputs(), golden_matrices()
for key in ("msc00726_csr", "2dcyl1_csr"):
    m = case(key)
    b = gm[key.split("_")[0] + "_b"]
    for prec in ("seqilu0", "sgs", "jacobi"):
        for solver in ("bicgstab", "gcr"):
            want = int(g[f"its_{key}_{prec}_{solver}"][0])
            view = bb.SRMatrixView(m)
            p = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
                prectype=SOLVER_TYPES[EXACT[prec]], bs=m.bs))
            p.compute()
            sol = bb.BiCGSTAB(view, p) if solver == "bicgstab" else bb.GCR(view, p, 30)
            sol.setParams(1e-10, 2000)
            x = np.zeros(m.dim)
            info = sol.solve(b, x)
            spread = ""
            if have_ref() and prec in ("jacobi", "seqilu0"):
                R = ref()
                its = []
                for nt in (1, 2, 3, 4, 6, 8, 12, 16):
                    R.set_num_threads(nt)
                    rp = R.prec(m, prec, nbuildsweeps=1, napplysweeps=1)
                    rp.compute()
                    _, it, rr, _ = R.solve(solver, rp, m, b, tol=1e-10, maxiter=2000, restart=30)
                    its.append(it)
                    rp.close()
                spread = f" reference over threads {its}"
            print(f"{key} {prec} {solver}: device {info.iters} golden {want} ({100.0*(info.iters-want)/want:+.1f} %){spread}", flush=True)
