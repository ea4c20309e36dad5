"""A/B timing of the asynchronous ILU(0) apply (development tool).  Usage: ab_apply.py [c2|c3s]"""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

for which in sys.argv[1:] or ["c2", "c3s"]:
    m = matgen.block_stencil((1024, 1024), 4, 1) if which == "c2" else matgen.block_stencil((96, 96, 96), 5, 2)
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=m.bs, nbuildsweeps=3, napplysweeps=3)
    p = bb.SRFactory().create_preconditioner(view, s)
    x = torch.randn(m.dim, dtype=torch.float64, device="cuda")
    z = torch.empty_like(x)
    p.compute()
    for _ in range(5):
        p.apply(x, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200
    e0.record()
    for _ in range(reps):
        p.apply(x, z)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/reps
    print(f"{which} TRI1={os.environ.get('B200_TRI1')} apply(3,3) {ms:.4f} ms  |z|={float(z.norm()):.12e}", flush=True)
