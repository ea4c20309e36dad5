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


def kernel_bytes(bs, N, nnz, c, scale=False):
    """Algorithmic bytes per launch of every kernel class; c = pattern_stats()."""
    b, b2 = bs, bs*bs
    out = {"spmv": (8*b2+4)*nnz + 4*(N+1) + 16*b*N}
    if b == 1:
        out["factor_lower"] = 32*c["nlower"] + 8*N + 24*c["npos_l"] + (8*c["nlower"] if scale else 0)
        out["factor_upper"] = 32*c["nuwork"] + 24*c["npos_u"]
        out["tri_lower"] = 12*c["nlower"] + 4*N + 24*N
        out["tri_upper"] = 12*(c["nupper"] - N) + 4*N + 32*N
    else:
        out["factor_lower"] = c["nlower"]*(16*b2+16) + 8*b2*N + c["npos_l"]*(16*b2+8)
        out["factor_upper"] = c["nuwork"]*(16*b2+16) + c["npos_u"]*(16*b2+8) + 8*b2*N
        out["tri_lower"] = c["nlower"]*(8*b2+4) + 8*N + 24*b*N
        out["tri_upper"] = (c["nupper"]-N)*(8*b2+4) + 8*N + 24*b*N + 8*b2*N
    return out


def measure(view, bs, N, nnz, nb=3, na=3, scale=False, steps=5, peak=None):
    """CUDA-event time of every kernel class on a resident matrix -> {kernel: {ms, bytes, gbs, frac}},
    plus the preconditioner (for further use), first-compute time and the pattern statistics."""
    peak = peak or PEAK
    x = torch.randn(N*bs, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["ilu0"], bs=bs, nbuildsweeps=nb, napplysweeps=na, scale=scale)
    p = bb.SRFactory().create_preconditioner(view, s)
    torch.cuda.synchronize(); t0 = time.time(); p.compute(); torch.cuda.synchronize(); setup = time.time() - t0
    c = p.pattern_stats()
    kb = kernel_bytes(bs, N, nnz, c, scale)
    p.compute(); p.apply(x, y)
    sf.profile_reset(); sf.profile_enable(True)
    for _ in range(steps):
        p.compute(); p.apply(x, y)
    prof = sf.profile_get(); sf.profile_enable(False)
    ms = {k: v[0]/max(v[1], 1) for k, v in prof.items()}
    ms["spmv"] = timeit(lambda: view.apply(x, y))
    out = {}
    for k, nbytes in kb.items():
        if ms.get(k, 0) > 0:
            gbs = nbytes/ms[k]/1e6
            out[k] = {"ms": ms[k], "bytes": int(nbytes), "gbs": gbs, "frac": gbs/peak}
    return out, p, setup, c, (x, y)


def report(title, m, nb=3, na=3, scale=False, sgs=False, levels=False, solve=None, view=None, dims=None):
    """m: host SRMatrix, or None with view = device-assembled SRMatrixView and dims = (bs, N, nnz)."""
    if m is not None:
        view = bb.SRMatrixView(m)
        b, N, nnz = m.bs, m.nbrows, m.nnzb
    else:
        b, N, nnz = dims
    b2 = b*b
    print(f"\n### {title}\n\nN = {N} (block) rows, nnzb = {nnz}, bs = {b}, scale = {scale}; peak = {PEAK:.0f} GB/s (measured)\n")
    print("| kernel | ms | algorithmic GB | GB/s | frac of measured peak |\n|---|---|---|---|---|")
    kern, p, setup, c, (x, y) = measure(view, b, N, nnz, nb, na, scale)
    names = {"spmv": "SpMV y=Ax", "factor_lower": "ILU(0) factor sweep, lower launch",
             "factor_upper": "ILU(0) factor sweep, upper launch", "tri_lower": "async L sweep (apply)",
             "tri_upper": "async U sweep (apply)"}
    for k, nm in names.items():
        if k in kern:
            row(nm, kern[k]["ms"], kern[k]["bytes"])
    tot_c = timeit(lambda: p.compute(), reps=5)
    tot_a = timeit(lambda: p.apply(x, y), reps=5)
    print(f"\ncompute() with {nb} sweeps: {tot_c:.3f} ms; apply() with {na} sweep pairs: {tot_a:.3f} ms; "
          f"first compute incl. device pattern build: {setup*1e3:.0f} ms; npos = {c['npos_l'] + c['npos_u']}")
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
        if isinstance(sgs, tuple):
            # SGS-preconditioned FGMRES time to solve (BASELINE config 3 as named)
            ps.set_sweeps(1, sgs[0])
            bvec = view.apply(torch.ones(N*b, dtype=torch.float64, device="cuda"))
            sol = bb.FGMRES(view, ps, 30); sol.setParams(1e-8, 2000)
            for rep in range(2):
                xs = torch.zeros_like(bvec)
                info = sol.solve(bvec, xs)
            print(f"\nFGMRES(30) + async block SGS ({sgs[0]} fwd + {sgs[0]} bwd sweeps): {info.iters} iterations, "
                  f"{info.walltime*1e3:.1f} ms to rel. residual {info.resnorm/info.bnorm:.1e}, max error {float((xs-1).abs().max()):.1e}")
    if levels:
        for mode, nm in ((0, "DAG wavefronts"),):
            pl = bb.SRFactory().create_preconditioner(view, bb.AsyncSolverSettings(
                prectype=SOLVER_TYPES["async_level_ilu0"], bs=b, nbuildsweeps=nb, scale=scale, level_mode=mode))
            torch.cuda.synchronize(); t0 = time.time(); pl.compute(); torch.cuda.synchronize(); ts = time.time() - t0
            nlev = pl.nlevels()
            t = timeit(lambda: pl.apply(x, y), reps=3, warm=1)
            print(f"\nlevel-scheduled exact ILU(0) apply ({nm}): {nlev} levels, {t:.3f} ms per apply "
                  f"(vs {tot_a/na:.3f} ms per async sweep pair); first compute incl. level build {ts*1e3:.0f} ms")
    if solve:
        p.set_sweeps(*solve)
        p.compute()
        bvec = view.apply(torch.ones(N*b, dtype=torch.float64, device="cuda"))
        for name, cls in (("FGMRES(30)", lambda: bb.FGMRES(view, p, 30)), ("GCR(30)", lambda: bb.GCR(view, p, 30))):
            sol = cls(); sol.setParams(1e-8, 2000)
            xs = torch.zeros_like(bvec)
            info = sol.solve(bvec, xs)
            print(f"\n{name} + async ILU(0) sweeps {solve}: {info.iters} iterations, {info.walltime*1e3:.1f} ms, "
                  f"rel. residual {info.resnorm/info.bnorm:.1e}, max error {float((xs-1).abs().max()):.1e}")


def device_view(kind, dims, bs=1, seed=0):
    """Device-assembled operators of the full BASELINE sizes -> (view, (bs, N, nnz))."""
    from blasted_b200 import matgen_device as md
    if kind == "block":
        nb, bp, bc, v = md.block_stencil_device(dims, bs, seed)
    else:
        bp, bc, v = md.poisson3d_device(dims, kind)
        nb = bp.numel() - 1
    view = bb.SRMatrixView.from_device(nb, bs, bp, bc, v)
    return view, (bs, nb, int(bc.numel()))


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
        v, d = device_view(7, (256, 256, 256))
        report("C1 - 7-point Poisson 256^3, CSR", None, solve=(5, 5), levels=True, sgs=True, view=v, dims=d)
        del v
    if "c2" in which:
        report("C2 - BSR bs=4, 1024x1024 cells (headline)", matgen.block_stencil((1024, 1024), 4, 20261020), sgs=(3,), solve=(3, 3))
    if "c3" in which:
        v, d = device_view("block", (128, 128, 128), 5, 20261021)
        report("C3 - BSR bs=5, 128^3 cells", None, sgs=(3,), solve=(3, 3), view=v, dims=d)
        del v
    if "c4" in which:
        v, d = device_view(27, (256, 256, 256))
        report("C4 - 27-point Poisson 256^3, CSR (scaled)", None, scale=True, levels=True, sgs=True, solve=(10, 20), view=v, dims=d)
        del v
    if "frontend" in which:
        frontend_report()
