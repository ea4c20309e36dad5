"""FGMRES(30) + async ILU(0) (5,5) on 7-point Poisson n^3, a bounded number of iterations
(development tool: run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
its = int(sys.argv[2]) if len(sys.argv) > 2 else 90
m = matgen.poisson3d(n)
A = bb.SRMatrixView(m)
s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=1, nbuildsweeps=5, napplysweeps=5)
p = bb.SRFactory().create_preconditioner(A, s)
p.compute()
b = A.apply(torch.ones(m.dim, dtype=torch.float64, device="cuda"))
for rep in range(2):
    x = torch.zeros_like(b)
    sol = bb.FGMRES(A, p, 30)
    sol.setParams(1e-30, its)
    info = sol.solve(b, x)
    print(f"n={n} its={info.iters} device {info.walltime*1e3:.2f} ms -> {info.walltime*1e3/info.iters:.3f} ms/iteration", flush=True)
