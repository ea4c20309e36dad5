"""Development probe (not part of the product): per-kernel device timings on the BASELINE shapes."""
import sys, os, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.solverfactory import SOLVER_TYPES

PEAK = 6550.1


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def run(name, m, nb=4, na=4, scale=False):
    b, N, nnz = m.bs, m.nbrows, m.nnzb
    t0 = time.time()
    view = bb.SRMatrixView(m)
    x = torch.randn(m.dim, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    med, mn = timeit(lambda: view.apply(x, y))
    bytes_spmv = (8*b*b+4)*nnz + 4*(N+1) + 16*b*N
    print(f"{name}: N={N} nnzb={nnz} bs={b} upload {time.time()-t0:.1f}s")
    print(f"  spmv        {med:8.3f} ms  {bytes_spmv/med/1e6:8.1f} GB/s  frac {bytes_spmv/med/1e6/PEAK:.2f}")
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=b, nbuildsweeps=nb, napplysweeps=na,
                               scale=scale)
    p = bb.SRFactory().create_preconditioner(view, s)
    t0 = time.time(); p.compute(); torch.cuda.synchronize()
    print(f"  first compute (pattern etc.) {time.time()-t0:.2f}s")
    npos = len(p.ilu_positions()[1])
    med, mn = timeit(lambda: p.compute(), reps=5)
    p.set_sweeps(0, na)
    med0, _ = timeit(lambda: p.compute(), reps=5)
    p.set_sweeps(nb, na)
    per = (med - med0)/nb
    bytes_f = nnz*(24*b*b+8) + 8*npos + 8*N + (8*b*N if scale else 0)
    print(f"  factor      {med:8.3f} ms total ({nb} sweeps; init+inv {med0:.3f}); per sweep {per:.3f} ms "
          f"{bytes_f/per/1e6:8.1f} GB/s frac {bytes_f/per/1e6/PEAK:.2f}  npos={npos}")
    z = torch.empty_like(x)
    med, mn = timeit(lambda: p.apply(x, z))
    p.set_sweeps(nb, 2*na)
    med2, _ = timeit(lambda: p.apply(x, z))
    per = (med2 - med)/na
    bytes_a = (8*b*b+4)*nnz + 12*N + 48*b*N
    print(f"  apply       {med:8.3f} ms total ({na} sweep pairs); per pair {per:.3f} ms "
          f"{bytes_a/per/1e6:8.1f} GB/s frac {bytes_a/per/1e6/PEAK:.2f}")
    for nsw in (1, 2, 3, 5, 10):
        p.set_sweeps(nsw, na)
        p.compute()
        print(f"    sweeps {nsw:2d}: rel nonlinear residual {p.ilu_residual():.3e}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c2", "c3s", "p128", "p27"]
    if "c2" in which:
        run("C2 bsr4 1024^2", matgen.block_stencil((1024, 1024), 4, 1))
    if "c3s" in which:
        run("C3-small bsr5 96^3", matgen.block_stencil((96, 96, 96), 5, 2))
    if "p128" in which:
        run("C1 poisson7 128^3", matgen.poisson3d(128))
    if "p256" in which:
        run("C1-large poisson7 256^3", matgen.poisson3d(256))
    if "p27" in which:
        run("C4-small poisson27 128^3", matgen.poisson3d(128, 27), scale=True)
