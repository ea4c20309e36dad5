"""Timing of the exact (level-ordered) ILU(0) apply: one-launch substitution vs per-level launches
replayed as a CUDA graph (B200_LEVEL_GRAPH=1).  Development tool.  Usage: ab_exact.py [c1|c2|c3s|c4]"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

for which in sys.argv[1:] or ["c1"]:
    m = {"c1": lambda: matgen.poisson3d(256), "c4": lambda: matgen.poisson3d(160, 27),
         "c2": lambda: matgen.block_stencil((1024, 1024), 4, 1),
         "c3s": lambda: matgen.block_stencil((96, 96, 96), 5, 2)}[which]()
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["sapilu0"], bs=m.bs, nbuildsweeps=5)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    torch.manual_seed(0)
    x = torch.randn(m.dim, dtype=torch.float64, device="cuda")
    z = torch.empty_like(x)
    for _ in range(3):
        p.apply(x, z)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        p.apply(x, z)
    e1.record(); torch.cuda.synchronize()
    nl = len(p.levels()[0]) - 1 if isinstance(p.levels(), tuple) else len(p.levels()) - 1
    print(f"{which} graph={os.environ.get('B200_LEVEL_GRAPH')} levels={nl} exact apply {e0.elapsed_time(e1)/reps:.3f} ms "
          f"|z|={float(z.norm()):.14e}", flush=True)
