"""FGMRES(30) + async ILU(0) (5,5), 7-point Poisson n^3 on one GPU: time to solve (development tool).
Run once plain and once with B200_FGMRES_SYNC=1 (host hand-over every iteration, the round-1 form)."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import blasted_b200 as bb
from blasted_b200.dist import Comm, DistMatrix, poisson3d_slab_device
from blasted_b200.solverfactory import SOLVER_TYPES

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
part, view = poisson3d_slab_device(n, 0, 1)
A = DistMatrix(Comm.single(), part, view)
s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=1, nbuildsweeps=5, napplysweeps=5)
p = bb.SRFactory().create_preconditioner(A.diag, s)
p.compute()
b = A.apply(torch.ones(A.local_dim(), dtype=torch.float64, device="cuda"))
for rep in range(2):
    x = torch.zeros_like(b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info = A.solve("fgmres", p, b, x, tol=1e-8, maxiter=6000, restart=30)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"n={n} sync={os.environ.get('B200_FGMRES_SYNC', '0')} its={info.iters} conv={info.converged} "
          f"device {info.walltime*1e3:.1f} ms wall {(t1-t0)*1e3:.1f} ms -> {info.walltime*1e3/info.iters:.3f} ms/it "
          f"err {float((x-1).abs().max()):.2e}", flush=True)
