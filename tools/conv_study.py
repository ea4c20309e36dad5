"""Development study: Krylov iterations vs async sweep counts on the reference fixtures."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import numpy as np
import blasted_b200 as bb
from blasted_b200.solverfactory import SOLVER_TYPES
from util import case, golden_matrices, golden_outputs

gm, g = golden_matrices(), golden_outputs()
for key in sys.argv[1:] or ["msc00726_csr", "2dcyl1_bsr4"]:
    m = case(key)
    b = gm[key.split("_")[0] + "_b"]
    want = int(g[f"its_{key}_seqilu0_bicgstab"][0])
    print(key, "reference sequential ILU0+BiCGSTAB its:", want)
    view = bb.SRMatrixView(m)
    for scale in (False, True):
        for nb, na in [(3, 3), (5, 5), (10, 10), (10, 30), (30, 60), (30, 150), (60, 300), (60, 600), (150, 1200)]:
            s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=m.bs, nbuildsweeps=nb,
                                       napplysweeps=na, scale=scale)
            p = bb.SRFactory().create_preconditioner(view, s)
            p.compute()
            sol = bb.BiCGSTAB(view, p)
            sol.setParams(1e-10, 1000)
            x = np.zeros(m.dim)
            info = sol.solve(b, x)
            print(f"  scale={scale} sweeps=({nb},{na}) its={info.iters} relres={info.resnorm/info.bnorm:.2e} ilu_res={p.ilu_residual():.2e}")
