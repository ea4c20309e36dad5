"""Profiling driver (development tool): block-ILU(0) factor + apply + SpMV, few launches."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
if which == "c2":
    m = matgen.block_stencil((1024, 1024), 4, 1)
elif which == "c3s":
    m = matgen.block_stencil((96, 96, 96), 5, 2)
elif which == "p128":
    m = matgen.poisson3d(128)
else:
    m = matgen.poisson3d(96, 27)
view = bb.SRMatrixView(m)
s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=m.bs, nbuildsweeps=2, napplysweeps=2)
p = bb.SRFactory().create_preconditioner(view, s)
x = torch.randn(m.dim, dtype=torch.float64, device="cuda")
z = torch.empty_like(x)
for _ in range(2):
    p.compute()
    p.apply(x, z)
    view.apply(x, z)
torch.cuda.synchronize()
print("done", bb.kernel_launches())
