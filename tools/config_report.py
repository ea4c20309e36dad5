"""Measured per-kernel throughput on the BASELINE.json configs (development/report tool).

    python tools/config_report.py [c1] [c2] [c3] [c4] > profiles/configs_rNN.md

Times every kernel class with CUDA events (median of repeats after warm-up), converts to achieved
GB/s with this implementation's algorithmic bytes and to a fraction of the
measured HBM peak (MEASURED_PEAKS.json)."""
import sys, os, json, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import blasted_b200 as bb
from blasted_b200 import matgen, solverfactory as sf
from blasted_b200.solverfactory import SOLVER_TYPES

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def row(name, ms, nbytes):
    if ms <= 0:
        print(f"| {name} | - | - | - | - |")
        return
    gbs = nbytes/ms/1e6
    print(f"| {name} | {ms:.3f} | {nbytes/1e9:.3f} | {gbs:.0f} | {gbs/PEAK:.2f} |")


def pattern_counts(m, posptr):
    rows = np.repeat(np.arange(m.nbrows), np.diff(m.browptr))
    cnt = np.diff(posptr)
    low = m.bcolind < rows
    dg = m.bcolind == rows
    return dict(nl=int(low.sum()), nu=int((~low).sum()), nuw=int(((~low) & ((cnt > 0) | dg)).sum()),
                pl=int(cnt[low].sum()), pu=int(cnt[~low].sum()))


def report(title, m, nb=3, na=3, scale=False, sgs=False, levels=False, solve=None):
    b, N, nnz = m.bs, m.nbrows, m.nnzb
    b2 = b*b
    print(f"\n### {title}\n\nN = {N} (block) rows, nnzb = {nnz}, bs = {b}, scale = {scale}; peak = {PEAK:.0f} GB/s (measured)\n")
    print("| kernel | ms | algorithmic GB | GB/s | frac of measured peak |\n|---|---|---|---|---|")
    view = bb.SRMatrixView(m)
    x = torch.randn(m.dim, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    row("SpMV y=Ax", timeit(lambda: view.apply(x, y)), (8*b2+4)*nnz + 4*(N+1) + 16*b*N)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=b, nbuildsweeps=nb, napplysweeps=na, scale=scale)
    p = bb.SRFactory().create_preconditioner(view, s)
    t0 = time.time(); p.compute(); torch.cuda.synchronize(); setup = time.time() - t0
    posptr, lowerp, _ = p.ilu_positions()
    c = pattern_counts(m, posptr)
    sf.profile_reset(); sf.profile_enable(True)
    for _ in range(5):
        p.compute(); p.apply(x, y)
    prof = sf.profile_get(); sf.profile_enable(False)
    avg = {k: v[0]/max(v[1], 1) for k, v in prof.items()}
    if b == 1:
        fl = 32*c["nl"] + 8*N + 24*c["pl"] + (8*c["nl"] if scale else 0)
        fu = 32*c["nuw"] + 24*c["pu"]
        tl = 12*c["nl"] + 4*N + 24*N
        tu = 12*(c["nu"] - N) + 4*N + 32*N
    else:
        fl = c["nl"]*(16*b2+16) + 8*b2*N + c["pl"]*(16*b2+8)
        fu = c["nuw"]*(16*b2+16) + c["pu"]*(16*b2+8) + 8*b2*N
        tl = c["nl"]*(8*b2+4) + 8*N + 24*b*N
        tu = (c["nu"]-N)*(8*b2+4) + 8*N + 24*b*N + 8*b2*N
    row("ILU(0) factor sweep, lower launch", avg["factor_lower"], fl)
    row("ILU(0) factor sweep, upper launch", avg["factor_upper"], fu)
    row("async L sweep (apply)", avg["tri_lower"], tl)
    row("async U sweep (apply)", avg["tri_upper"], tu)
    tot_c = timeit(lambda: p.compute(), reps=5)
    tot_a = timeit(lambda: p.apply(x, y), reps=5)
    print(f"\ncompute() with {nb} sweeps: {tot_c:.3f} ms; apply() with {na} sweep pairs: {tot_a:.3f} ms; "
          f"first compute incl. device pattern build: {setup*1e3:.0f} ms; npos = {len(lowerp)}")
    res = []
    for k in (1, 2, 3, 5, 10, 20):
        p.set_sweeps(k, na); info = None
        p.compute(); res.append(f"{k}: {p.ilu_residual():.2e}")
    print("\nnonlinear residual sum|(A-LU)_S| after k sweeps: " + ", ".join(res))
    if sgs:
        ps = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
            prectype=SOLVER_TYPES["sgs"], bs=b, napplysweeps=na))
        ps.compute()
        t = timeit(lambda: ps.apply(x, y), reps=5)
        print(f"\nasync SGS apply ({na} fwd + {na} bwd sweeps): {t:.3f} ms "
              f"({na*((8*b2+4)*nnz + 12*N + 48*b*N + 8*b2*N)/t/1e6:.0f} GB/s)")
    if levels:
        for mode, nm in ((0, "DAG wavefronts"),):
            pl = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
                prectype=SOLVER_TYPES["async_level_ilu0"], bs=b, nbuildsweeps=nb, scale=scale, level_mode=mode))
            t0 = time.time(); pl.compute(); torch.cuda.synchronize(); ts = time.time() - t0
            ptr, _ = pl.levels()
            t = timeit(lambda: pl.apply(x, y), reps=3, warm=1)
            print(f"\nlevel-scheduled exact ILU(0) apply ({nm}): {len(ptr)-1} levels, {t:.3f} ms per apply "
                  f"(vs {tot_a/na:.3f} ms per async sweep pair); level build {ts*1e3:.0f} ms")
    if solve:
        p.set_sweeps(*solve)
        p.compute()
        bvec = view.apply(torch.ones(m.dim, dtype=torch.float64, device="cuda"))
        for name, cls in (("FGMRES(30)", lambda: bb.FGMRES(view, p, 30)), ("GCR(30)", lambda: bb.GCR(view, p, 30))):
            sol = cls(); sol.setParams(1e-8, 2000)
            xs = torch.zeros_like(bvec)
            info = sol.solve(bvec, xs)
            print(f"\n{name} + async ILU(0) sweeps {solve}: {info.iters} iterations, {info.walltime*1e3:.1f} ms, "
                  f"rel. residual {info.resnorm/info.bnorm:.1e}, max error {float((xs-1).abs().max()):.1e}")


def frontend_report():
    """Front end: device conversion / reordering times on the C2 matrix, with
    the CPU restatement of the reference algorithm (oracle, 1 thread) on a 1/16 sample beside it."""
    import ctypes as C
    from blasted_b200._lib import lib, check
    from blasted_b200.frontend import Reordering, FORWARD, INVERSE
    m = matgen.block_stencil((1024, 1024), 4, 20261020)
    nb = m.nbrows
    rows = torch.repeat_interleave(torch.arange(nb, device="cuda", dtype=torch.int32),
                                   torch.as_tensor(np.diff(m.browptr), device="cuda"))
    cols = torch.as_tensor(m.bcolind, device="cuda")
    k = torch.arange(16, device="cuda", dtype=torch.int32)
    r = (rows[:, None]*4 + (k % 4)[None, :]).reshape(-1)
    c = (cols[:, None]*4 + (k // 4)[None, :]).reshape(-1)
    v = torch.as_tensor(m.vals, device="cuda")
    p = torch.randperm(r.numel(), device="cuda")
    r, c, v = r[p].contiguous(), c[p].contiguous(), v[p].contiguous()
    del p, rows, cols
    nnz = r.numel()
    print("\n### Front end - coordinate input and reordering on the C2 matrix\n")
    ts = []
    for _ in range(4):
        h = C.c_void_p()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        check(lib.b200_mat_create_coo(m.dim, nnz, C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()),
                                      C.c_void_p(v.data_ptr()), 4, 0, 1, C.byref(h)))
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        if _ < 3:
            lib.b200_mat_destroy(h)
    t_conv = min(ts[1:])
    print(f"{nnz/1e6:.1f} M scrambled scalar triplets (device-resident) -> BSR4, {nb} block rows, {m.nnzb} blocks: "
          f"{t_conv*1e3:.1f} ms ({nnz*16/t_conv/1e9:.0f} GB/s of triplets)")
    perm = torch.randperm(nb, device="cuda").to(torch.int32)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        check(lib.b200_mat_reorder(h, C.c_void_p(perm.data_ptr()), C.c_void_p(perm.data_ptr()), 0, 1))
        check(lib.b200_mat_reorder(h, C.c_void_p(perm.data_ptr()), C.c_void_p(perm.data_ptr()), 1, 1))
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0)/2)
    print(f"\nsymmetric random permutation of the resident matrix (rows + columns, re-sorted, diagonals "
          f"located again): {min(ts)*1e3:.1f} ms per application ({m.nnzb*128*2/min(ts)/1e9:.0f} GB/s of blocks moved)")
    lib.b200_mat_destroy(h)
    try:
        from oracle import orc
        sub = matgen.block_stencil((256, 256), 4, 20261020)
        sp = sub.to_scipy().tocoo()
        q = np.random.default_rng(0).permutation(sp.nnz)
        rr, cc, vv = sp.row[q].astype(np.int32), sp.col[q].astype(np.int32), sp.data[q]
        t0 = time.perf_counter(); orc().coo_convert(sub.dim, rr, cc, vv, 4, False); t1 = time.perf_counter() - t0
        pp = np.random.default_rng(1).permutation(sub.nbrows).astype(np.int32)
        t0 = time.perf_counter(); orc().reorder_matrix(sub, pp, pp, False); t2 = time.perf_counter() - t0
        print(f"\nCPU restatement of the reference algorithm (oracle, 1 thread) on the 256x256-cell sample "
              f"(1/16 of the work): conversion {t1*1e3:.0f} ms, permutation {t2*1e3:.0f} ms")
    except Exception as e:                              # noqa: BLE001
        print(f"\n(oracle not available for the CPU comparison: {e})")


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "frontend"]
    print(f"# Per-kernel throughput on the BASELINE configs ({torch.cuda.get_device_name(0)})")
    if "c1" in which:
        report("C1 - 7-point Poisson 256^3, CSR", matgen.poisson3d(256), solve=(5, 5), levels=True, sgs=True)
    if "c2" in which:
        report("C2 - BSR bs=4, 1024x1024 cells (headline)", matgen.block_stencil((1024, 1024), 4, 20261020), sgs=True, solve=(3, 3))
    if "c3" in which:
        report("C3 - BSR bs=5, 128^3 cells", matgen.block_stencil((128, 128, 128), 5, 20261021), sgs=True, solve=(3, 3))
    if "c4" in which:
        report("C4 - 27-point Poisson 192^3, CSR (scaled)", matgen.poisson3d(192, 27), scale=True, levels=True, sgs=True, solve=(10, 20))
    if "frontend" in which:
        frontend_report()
