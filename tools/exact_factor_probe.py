"""How long the exact ('sequential') factorisation takes: sweeps to the bitwise fixed point.
Development tool.  Usage: exact_factor_probe.py [c1|c2|c3s|c4]"""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

for which in sys.argv[1:] or ["c1"]:
    m = {"c1": lambda: matgen.poisson3d(256), "c4": lambda: matgen.poisson3d(160, 27),
         "p128": lambda: matgen.poisson3d(128),
         "c2": lambda: matgen.block_stencil((1024, 1024), 4, 1),
         "c3s": lambda: matgen.block_stencil((96, 96, 96), 5, 2)}[which]()
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["sfilu0"], bs=m.bs, nbuildsweeps=1)
    p = bb.SRFactory().create_preconditioner(view, s)
    p.compute()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p.compute()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"{which}: exact factorisation {1e3*(t1-t0):.1f} ms, last_times={p.last_times()}", flush=True)
