"""Development study: time-to-solve vs sweep counts / preconditioner type (7-point Poisson n^3, GCR(30))."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = matgen.poisson3d(n)
view = bb.SRMatrixView(m)
b = view.apply(torch.ones(m.dim, dtype=torch.float64, device="cuda"))
cases = [("ilu0", 5, 5), ("ilu0", 5, 3), ("ilu0", 5, 4), ("ilu0", 5, 7), ("sgs", 0, 3), ("async_level_ilu0", 10, 1)]
for pt, nb, na in cases:
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES[pt], bs=1, nbuildsweeps=nb, napplysweeps=na)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    for restart, cls, nm in ((30, bb.GCR, "GCR"), (30, bb.FGMRES, "FGMRES"), (60, bb.FGMRES, "FGMRES")):
        sol = cls(view, p, restart); sol.setParams(1e-8, 3000)
        x = torch.zeros_like(b)
        info = sol.solve(b, x)
        print(f"{pt:18s} sweeps=({nb},{na}) {nm}({restart}): its={info.iters:5d} time={info.walltime*1e3:9.1f} ms ms/it={info.walltime*1e3/max(info.iters,1):6.2f} relres={info.resnorm/info.bnorm:.1e}", flush=True)
