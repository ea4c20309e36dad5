"""Profiling driver (development tool): ILU(0) factor + apply + SpMV on one BASELINE config at its
full size, few launches.  Usage: prof_c2.py c1|c2|c3|c4|c3s|p128"""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.dirname(__file__)))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES
import config_report as cr

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
scale = False
if which == "c1":
    view, (bs, N, nnz) = cr.device_view(7, (256, 256, 256))
elif which == "c2":
    view, (bs, N, nnz) = cr.device_view("block", (1024, 1024), 4, 1)
elif which == "c3":
    view, (bs, N, nnz) = cr.device_view("block", (128, 128, 128), 5, 2)
elif which == "c4":
    view, (bs, N, nnz) = cr.device_view(27, (256, 256, 256))
    scale = True
elif which == "c3s":
    view, (bs, N, nnz) = cr.device_view("block", (96, 96, 96), 5, 2)
else:
    view, (bs, N, nnz) = cr.device_view(7, (128, 128, 128))
s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=bs, nbuildsweeps=2, napplysweeps=2, scale=scale)
p = bb.SRFactory().create_preconditioner(view, s)
x = torch.randn(N*bs, dtype=torch.float64, device="cuda")
z = torch.empty_like(x)
for _ in range(2):
    p.compute()
    p.apply(x, z)
    view.apply(x, z)
torch.cuda.synchronize()
c = p.pattern_stats()
kb = cr.kernel_bytes(bs, N, nnz, c, scale)
print("algorithmic_bytes", which, " ".join(f"{k}={v}" for k, v in kb.items()))
print("done", bb.kernel_launches())
