"""Development study: FGMRES(30) time-to-solve on the bench's 7-point n^3 problem (one GPU,
device-assembled, operator as in bench.run_fgmres) against the sweep counts of the async ILU(0).
    python tools/solve_study512.py [n] [nb,na ...]"""
import os
import sys
import time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import blasted_b200 as bb
from blasted_b200 import solverfactory as sf
from blasted_b200.dist import Comm, DistMatrix, poisson3d_slab_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
# each case: nb,na[,type]
cases = [tuple(a.split(",")) for a in sys.argv[2:]] or [("5", "5"), ("5", "4"), ("5", "3")]
part, view = poisson3d_slab_device(n, 0, 1)
A = DistMatrix(Comm.single(), part, view)
b = A.apply(torch.ones(A.local_dim(), dtype=torch.float64, device="cuda"))
for c in cases:
    nb, na, pt = int(c[0]), int(c[1]), (c[2] if len(c) > 2 else "ilu0")
    s = bb.AsyncSolverSettings(prectype=sf.SOLVER_TYPES[pt], bs=1, nbuildsweeps=nb, napplysweeps=na)
    prec = bb.SRFactory().create_preconditioner(A.diag, s)
    prec.compute()
    x = torch.zeros_like(b)
    A.solve("fgmres", prec, b, x, tol=1e-8, maxiter=31, restart=30)
    x.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prec.compute()
    info = A.solve("fgmres", prec, b, x, tol=1e-8, maxiter=6000, restart=30)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1)
    print(f"n={n} {pt} sweeps=({nb},{na}) its={info.iters} converged={info.converged} time={t:.0f} ms "
          f"ms/it={t/max(info.iters,1):.2f} err={float((x-1).abs().max()):.2e}", flush=True)
    del prec
