"""A/B timing of the asynchronous ILU(0) factor sweeps (development tool).
Usage: [B200_LIB=...] ab_factor.py [c1|c4|c2|c3s ...]"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen, solverfactory as sf
from blasted_b200.solverfactory import SOLVER_TYPES

for which in sys.argv[1:] or ["c1"]:
    m = {"c1": lambda: matgen.poisson3d(256), "c4": lambda: matgen.poisson3d(160, 27),
         "c2": lambda: matgen.block_stencil((1024, 1024), 4, 1),
         "c3s": lambda: matgen.block_stencil((96, 96, 96), 5, 2),
         "c3": lambda: matgen.block_stencil((128, 128, 128), 5, 2),
         "b27": lambda: matgen.csr_to_bsr(matgen.poisson3d(0, 27, dims=(4*40, 40, 40)), 4)}[which]()
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=m.bs, nbuildsweeps=3, napplysweeps=1,
                               scale=(which == "c4"))
    p = bb.SRFactory().create_preconditioner(view, s)
    for _ in range(3):
        p.compute()
    sf.profile_enable(True); sf.profile_reset()
    x = torch.randn(m.dim, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    for _ in range(20):
        p.compute(); p.apply(x, y)
    torch.cuda.synchronize()
    prof = sf.profile_get()
    out = " ".join(f"{k}={v[0]/max(v[1],1):.4f}ms" for k, v in prof.items() if v[1])
    print(f"{which} LIB={os.path.basename(os.environ.get('B200_LIB', 'default'))} NO_PF={os.environ.get('B200_NO_PF','0')} {out}", flush=True)
