"""Development probe: BSR SpMV on the strict-lower part only (what a split-storage L sweep would stream)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import blasted_b200 as bb
from blasted_b200 import matgen
PEAK = 6550.1
def timeit(fn, reps=9, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for name, m in (("C2 bs4", matgen.block_stencil((1024, 1024), 4, 1)), ("C3s bs5", matgen.block_stencil((96, 96, 96), 5, 2))):
    rows = np.repeat(np.arange(m.nbrows), np.diff(m.browptr))
    b2 = m.bs*m.bs
    for part, mask in (("lower", m.bcolind < rows), ("upper", m.bcolind > rows), ("full", np.ones(m.nnzb, bool))):
        ptr = np.zeros(m.nbrows+1, np.int32); np.cumsum(np.bincount(rows[mask], minlength=m.nbrows), out=ptr[1:])
        sub = matgen.SRMatrix(m.nbrows, m.bs, ptr, m.bcolind[mask].copy(), m.vals.reshape(-1, b2)[mask].reshape(-1).copy(), None)
        v = bb.SRMatrixView(sub)
        x = torch.randn(m.dim, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
        t = timeit(lambda: v.apply(x, y))
        nbytes = (8*b2+4)*sub.nnzb + 4*(m.nbrows+1) + 16*m.bs*m.nbrows
        print(f"{name} {part:6s}: {t:.3f} ms  {nbytes/t/1e6:.0f} GB/s  frac {nbytes/t/1e6/PEAK:.2f}")
