#!/usr/bin/env python
"""bench.py - async block-ILU(0) factor + apply throughput on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref)

Workload (config.workload = "C2"): synthetic 2D compressible-flow-like BSR bs=4 Jacobian, 1024x1024
cells = 1 048 576 block rows, 5-point block stencil (BASELINE.json configs[1]).
One STEP = compute() [init + nbuild async block-ILU(0) sweeps + diagonal-block inversion] followed by
apply() [napply async lower + upper block-triangular sweeps], sweeps (3,3).
metric = algorithmic HBM bytes moved by one step / time ("async ILU factor+apply HBM GB/s").
At N > 1 every rank factors and applies its own subdomain of the same size (block-Jacobi: the
preconditioner needs no communication) -> weak scaling, value = bytes of all ranks / max time.
See DESIGN.md section "Measurement" for the byte formulas.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261018 + 2
NBUILD, NAPPLY = 3, 3
FGMRES_SWEEPS = (5, 3)      # fastest of tools/solve_study512.py (profiles/solve_study512_r02.log)


# ----------------------------------------------------------------------------- workload + bytes

def load_matgen():
    """blasted_b200/matgen.py is pure numpy; it is loaded by path so that the reference arm does not
    import the package (which would map libblasted_b200.so into that process)."""
    import importlib.util
    if "b200_matgen" in sys.modules:
        return sys.modules["b200_matgen"]
    spec = importlib.util.spec_from_file_location(
        "b200_matgen", os.path.join(ROOT, "blasted_b200", "matgen.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b200_matgen"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_matrix(rank, cells):
    return load_matgen().block_stencil((cells, cells), 4, SEED + 1000*rank)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


REF_KIND = "reference"     # the driver's vocabulary; what it is exactly goes into `build`
REF_BUILD = ("unmodified reference sources (OpenMP) compiled against the Eigen/Boost header shim of "
             "oracle/shim, g++ -O3 -march=x86-64-v3")


def reference_bytes(m, npos, nbuild, napply):
    """Compulsory DRAM traffic of one step of the REFERENCE algorithm (SURVEY.md section 8(d);
    int32 = 4 B, fp64 = 8 B): every stored entry is recomputed in every sweep."""
    b, N, nnz = m.bs, m.nbrows, m.nnzb
    b2 = b*b
    sweep = nnz*(24*b2 + 8) + 8*npos + 8*N              # read A, read+write factor, pattern
    init = nnz*16*b2                                     # read A, write factor
    dinv = 16*b2*N + 4*N                                 # diagonal block inversion
    apply_pair = (8*b2 + 4)*nnz + 12*N + 48*b*N          # factor + indices once, vectors
    return {"step": init + nbuild*sweep + dinv + napply*apply_pair}


def kernel_bytes(m, pat):
    """Algorithmic bytes of ONE launch of each kernel class of THIS implementation.

    pat: nlower, nupper, nuwork (upper entries with products or on the diagonal; the others are
    U_ij = A_ij, set by the initialisation and not re-touched), npos_l / npos_u (products of lower /
    upper entries).  Per launch every distinct array element is counted once."""
    b, N = m.bs, m.nbrows
    b2 = b*b
    # lower launch: {entry,col} list 8 B, A block read, factor block written, posptr 8 B,
    # U_jj^-1 read once per block row, two partner blocks + 8 B pair per product
    lower = pat["nlower"]*(16*b2 + 16) + 8*b2*N + pat["npos_l"]*(16*b2 + 8)
    # upper launch: 16 B work-list item, A block read, factor block written per work entry, two
    # partner blocks + pair per product, refreshed inverse written per block row
    upper = pat["nuwork"]*(16*b2 + 16) + pat["npos_u"]*(16*b2 + 8) + 8*b2*N
    # triangular sweeps: blocks + column indices of their half, browptr/diagind, rhs read,
    # solution gathered and written; the upper sweep also reads U_ii^-1
    tri_l = pat["nlower"]*(8*b2 + 4) + 8*N + 24*b*N
    tri_u = (pat["nupper"] - N)*(8*b2 + 4) + 8*N + 24*b*N + 8*b2*N
    return {"factor_lower": lower, "factor_upper": upper, "tri_lower": tri_l, "tri_upper": tri_u}


def algorithmic_bytes(m, pat, nbuild, napply):
    """Compulsory HBM traffic of one step of THIS implementation."""
    b, N, nnz = m.bs, m.nbrows, m.nnzb
    b2 = b*b
    kb = kernel_bytes(m, pat)
    init = nnz*16*b2                                     # read A, write factor
    dinv0 = 16*b2*N + 4*N                                # inverses of the initial diagonal blocks
    sweep = kb["factor_lower"] + kb["factor_upper"]
    vec = 3*8*b*N                                        # y := 0, z := y (init_jacobi)
    apply_pair = kb["tri_lower"] + kb["tri_upper"]
    return {"factor_sweep": sweep, "init": init, "diag_invert": dinv0, "apply_pair": apply_pair,
            "step": init + dinv0 + nbuild*sweep + vec + napply*apply_pair}


# ----------------------------------------------------------------------------- clocks

class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons DURING the timed region: one long-running
    `nvidia-smi -lms 20` whose lines are collected until stop()."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2+i].lower().startswith("active")
                                                         for s in self.samples if len(s) > 2+i)]
        mx = self.samples[0][1]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(mx) if mx.replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


# ----------------------------------------------------------------------------- reference (CPU) arm

def cpu_step_time(m, steps, warmup, nbuild, napply):
    """Times the reference's own OpenMP path (oracle/_ref) or, failing that, the sequential oracle
    port, on the host cores.  Returns (seconds per step, kind, cores)."""
    from oracle import have_ref, ref, orc
    r = np.random.default_rng(SEED).standard_normal(m.dim)
    if have_ref():
        R = ref()
        # all host threads it can use (torchrun exports OMP_NUM_THREADS=1 to its workers)
        try:
            ncpu = len(os.sched_getaffinity(0))
        except Exception:
            ncpu = os.cpu_count() or 1
        R.set_num_threads(ncpu)
        cores = R.num_threads()
        p = R.prec(m, "ilu0", nbuildsweeps=nbuild, napplysweeps=napply, thread_chunk_size=128,
                   fact_init="init_original", apply_init="init_jacobi")
        p.compute()                       # first call builds the (serial) position lists: not timed
        ts = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            p.compute()
            p.apply(r)
            if it >= warmup:
                ts.append(time.perf_counter() - t0)
        p.close()
        return float(np.mean(ts)), REF_KIND, cores
    O = orc()
    plist = O.ilu_positions(m)
    ts = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ilu = O.ilu0_init(m, None, "init_original")
        O.ilu0_sweeps(m, plist, None, nbuild, ilu)
        O.ilu0_invert_diag(m, ilu)
        O.ilu0_apply(m, ilu, None, napply, "init_jacobi", r)
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)), "port", 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                   # rank 0 alone runs the CPU arm
    cells = args.cells
    m = make_matrix(0, cells)
    npos = 2*(cells*cells) - 2*cells                      # 5-point stencil: 2 products per diagonal
    by = reference_bytes(m, npos, NBUILD, NAPPLY)
    sec, kind, cores = cpu_step_time(m, args.steps, args.warmup, NBUILD, NAPPLY)
    val = by["step"]/sec/1e9
    line = {"impl": "reference", "metric": "async_ilu0_factor_apply_hbm_gbs", "value": val,
            "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec*1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(m, cells),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": kind,
                             "build": REF_BUILD if kind == REF_KIND else "oracle port (plain C, 1 thread)",
                             "cpu_model": cpu_model(),
                             "sample": f"full C2 step ({NBUILD} factor sweeps + {NAPPLY} apply sweep "
                                       f"pairs) x {args.steps} on the host cores"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_config(m, cells):
    return {"workload": "C2", "matrix": f"synthetic BSR bs=4 5-point block stencil, {cells}x{cells} cells",
            "block_rows": m.nbrows, "nnzb": m.nnzb, "prectype": "ilu0 (async block-ILU(0))",
            "nbuildsweeps": NBUILD, "napplysweeps": NAPPLY, "fact_init": "init_original",
            "apply_init": "init_jacobi", "parallelism": "one subdomain per GPU (block-Jacobi)",
            "l2_policy": "inputs larger than L2 (matrix+factor 1.3 GB vs 126 MB L2)"}


# ----------------------------------------------------------------------------- our arm

def run_b200(args):
    import torch
    import torch.distributed as dist
    import blasted_b200 as bb
    from blasted_b200 import solverfactory as sf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if bb.device_count() == 0:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells = args.cells
    m = make_matrix(rank, cells)
    view = bb.SRMatrixView(m)
    s = bb.AsyncSolverSettings(prectype=sf.SOLVER_TYPES["ilu0"], bs=4, nbuildsweeps=NBUILD,
                               napplysweeps=NAPPLY)
    prec = bb.SRFactory().create_preconditioner(view, s)
    prec.compute()                                        # pattern + first factorisation (setup)
    posptr, lowerp, _ = prec.ilu_positions()
    npos = len(lowerp)
    rows = np.repeat(np.arange(m.nbrows), np.diff(m.browptr))
    cnt = np.diff(posptr)
    islower = m.bcolind < rows
    isdiag = m.bcolind == rows
    pat = {"nlower": int(islower.sum()), "nupper": int((~islower).sum()),
           "nuwork": int(((~islower) & ((cnt > 0) | isdiag)).sum()),
           "npos_l": int(cnt[islower].sum()), "npos_u": int(cnt[~islower].sum())}
    by = algorithmic_bytes(m, pat, NBUILD, NAPPLY)
    kby = kernel_bytes(m, pat)
    ref_by = reference_bytes(m, npos, NBUILD, NAPPLY)

    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    r_dev = torch.randn(m.dim, dtype=torch.float64, device="cuda", generator=gen)
    z_dev = torch.empty_like(r_dev)

    def step():
        prec.compute()
        prec.apply(r_dev, z_dev)

    # ---- device-resident timing: W warm-up steps, then exactly K timed steps
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
        time.sleep(0.15)                                   # let the sampler come up
    sf.profile_reset()
    sf.profile_enable(True)
    bb.reset_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = bb.kernel_launches()
    prof = sf.profile_get()
    sf.profile_enable(False)
    sampler.stop()
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world*by["step"]*args.steps/(ms_max*1e-3)/1e9

    # ---- end to end through the host-pointer C ABI: new matrix values and r from pinned host
    # memory every step, z read back to the host every step
    vals_pin = torch.from_numpy(m.vals).pin_memory()
    r_pin = r_dev.cpu().pin_memory()
    z_pin = torch.empty(m.dim, dtype=torch.float64).pin_memory()
    vals_np, r_np, z_np = vals_pin.numpy(), r_pin.numpy(), z_pin.numpy()

    def step_e2e():
        prec.compute(vals_np)                             # H2D of the Jacobian values + factorisation
        prec.apply(r_np, z_np)                            # H2D r, kernels, D2H z

    esteps = max(1, min(args.steps, 5))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(esteps):
        step_e2e()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())/esteps
    e2e_val = world*by["step"]/e2e_s/1e9
    checksum = float(np.abs(z_np).sum())
    assert np.isfinite(checksum)
    # what the PCIe link alone allows: the same bytes as plain pinned copies (all ranks at once,
    # as in the step), measured here
    dev_buf = torch.empty(m.vals.size, dtype=torch.float64, device="cuda")
    zsrc = torch.empty(m.dim, dtype=torch.float64, device="cuda")
    link = {}
    for name, fn, nbytes in (("h2d", lambda: dev_buf.copy_(vals_pin, non_blocking=True), m.vals.nbytes),
                             ("d2h", lambda: z_pin.copy_(zsrc, non_blocking=True), z_np.nbytes)):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tt = torch.tensor([(time.perf_counter() - t0)/3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        link[name] = nbytes/float(tt.item())/1e9
    del dev_buf, zsrc
    link_bound_s = (m.vals.nbytes + r_np.nbytes)/(link["h2d"]*1e9) + z_np.nbytes/(link["d2h"]*1e9)

    # ---- second half of the metric: FGMRES (restarted GCR == FGMRES in exact arithmetic) time to
    # solve, 7-point Poisson n^3 row-partitioned into z-slabs over the ranks (STRONG scaling),
    # block-Jacobi async ILU(0), NCCL halo exchange + all-reduce
    # ---- the other BASELINE configs at their full sizes (N = 1 only): per-kernel rooflines of the
    # scalar (C1, C4) and bs=5 (C3) paths and of SpMV, so that the driver's record carries them
    configs = None
    if world == 1 and not args.no_configs:
        del prec, view
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        configs = run_configs(peak_hbm())

    fgmres, fgmres_ok = None, None
    if args.fgmres_n > 0:
        # free the headline workload first: the 512^3 operator, its factor and 61 basis vectors
        # take most of one GPU's HBM
        if configs is None:
            del prec, view
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            fgmres = run_fgmres(args.fgmres_n, rank, world, dist if world > 1 else None)
            fgmres_ok = bool(fgmres["converged"]) and fgmres["max_abs_error"] < 1e-3
        except Exception as e:                             # reported AND reflected in the exit code
            fgmres = {"error": str(e)[:300]}
            fgmres_ok = False

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the live per-launch CUDA-event times
    peak, peak_src = peak_hbm(), peak_hbm_source()
    kernels = {}
    for k in ("factor_lower", "factor_upper", "tri_lower", "tri_upper"):
        tot, cnt = prof[k]
        if cnt:
            avg = tot/cnt
            kernels[k] = {"avg_ms": avg, "launches": cnt, "share_of_step": tot/ms,
                          "achieved_gbs": kby[k]/(avg*1e-3)/1e9, "bytes_per_launch": kby[k]}
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"]) if kernels else None
    roofline = None
    if dom:
        ach = kernels[dom]["achieved_gbs"]
        traffic = None
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
        for tname in ("traffic_r02.json", "traffic_r01.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                t = json.load(open(tpath))
                t = t.get("C2", t).get(dom)
                traffic = t.get("dram_bytes_per_launch") if isinstance(t, dict) else t
                break
        roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach/peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": kby[dom]}

    # ---- CPU baseline beside it (rank 0, N=1 only), bounded sample
    cpu = None
    if world == 1 and not args.no_cpu:
        csteps = 2
        sec, kind, cores = cpu_step_time(m, csteps, 1, NBUILD, NAPPLY)
        cpu = {"value": ref_by["step"]/sec/1e9, "unit": "GB/s", "cores": cores, "kind": kind,
               "build": REF_BUILD if kind == REF_KIND else "oracle port (plain C, 1 thread)",
               "cpu_model": cpu_model(), "ms_per_step": sec*1e3,
               "sample": f"full C2 step x {csteps} (1 warm-up) on the host cores, same matrix"}

    line = {"metric": "async_ilu0_factor_apply_hbm_gbs", "value": value, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max/args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(m, cells),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_val, "unit": "GB/s",
                    "h2d_bytes_per_step": int(m.vals.nbytes + r_np.nbytes),
                    "d2h_bytes_per_step": int(z_np.nbytes), "steps": esteps,
                    "ms_per_step": e2e_s*1e3, "h2d_gbs_measured": link["h2d"], "d2h_gbs_measured": link["d2h"],
                    "link_bound_ms_per_step": link_bound_s*1e3,
                    "frac_of_link_bound": link_bound_s/e2e_s,
                    "note": "host-pointer C ABI: the 704 MB of new Jacobian values cross PCIe every step; "
                            "link_bound = the same bytes as plain pinned copies, device work excluded"},
            "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "configs": configs,
            "fgmres": fgmres, "fgmres_ok": fgmres_ok,
            "algorithmic_bytes_per_step": by["step"],
            "reference_algorithm_bytes_per_step": ref_by["step"],
            "frac_of_peak_whole_step": value/world/peak}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if fgmres_ok is False:
        sys.exit(3)                                        # a broken solve leg must not look green


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def peak_hbm_source():
    return ("measured (MEASURED_PEAKS.json hbm_gbs)" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
            else "fallback (B200_PROFILING.md)")


def run_configs(peak):
    """BASELINE.json configs other than the headline, each at its FULL size, assembled on the device:
    C1 7-point Poisson 256^3 CSR, C3 BSR bs=5 128^3 cells, C4 27-point Poisson 256^3 CSR (scaled
    factorisation).  Per kernel class: CUDA-event ms per launch, algorithmic bytes per launch
    (DESIGN.md section 3), achieved GB/s and the fraction of the measured HBM peak - SpMV included
    (north_star: BSR SpMV >= 60 % of HBM bandwidth).  C2's SpMV is reported with C3's."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config_report as cr
    out = {}
    specs = (("C1", "7-point Poisson 256^3 CSR", lambda: cr.device_view(7, (256, 256, 256)), dict(scale=False)),
             ("C2_spmv", "BSR bs=4 1024x1024 cells (SpMV only)", lambda: cr.device_view("block", (1024, 1024), 4, SEED), None),
             ("C3", "BSR bs=5 128^3 cells", lambda: cr.device_view("block", (128, 128, 128), 5, SEED + 1), dict(scale=False)),
             ("C4", "27-point Poisson 256^3 CSR, scaled factorisation", lambda: cr.device_view(27, (256, 256, 256)), dict(scale=True)))
    for key, desc, make, kw in specs:
        try:
            view, (bs, N, nnz) = make()
            if kw is None:
                x = torch.randn(N*bs, dtype=torch.float64, device="cuda")
                y = torch.empty_like(x)
                ms = cr.timeit(lambda: view.apply(x, y))
                nbytes = (8*bs*bs + 4)*nnz + 4*(N + 1) + 16*bs*N
                kern = {"spmv": {"ms": ms, "bytes": int(nbytes), "gbs": nbytes/ms/1e6, "frac": nbytes/ms/1e6/peak}}
                setup = None
            else:
                kern, p, setup, c, _ = cr.measure(view, bs, N, nnz, 3, 3, kw["scale"], steps=3, peak=peak)
                del p
            out[key] = {"matrix": desc, "block_rows": N, "nnzb": nnz, "bs": bs, "sweeps": [3, 3],
                        "first_compute_incl_pattern_build_ms": None if setup is None else setup*1e3,
                        "kernels": {k: {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()}
                                    for k, v in kern.items()}}
            del view
        except Exception as e:                              # noqa: BLE001 - reported in the line
            out[key] = {"matrix": desc, "error": str(e)[:300]}
        import gc
        gc.collect()
        torch.cuda.empty_cache()
    return out


def run_fgmres(n, rank, world, dist, reps=1):
    """Time to solve A x = b (7-point Poisson n^3, x* = 1) to rel. residual 1e-8 with FGMRES(30)
    preconditioned by per-subdomain async ILU(0) (FGMRES_SWEEPS build sweeps / apply sweep pairs;
    measured fastest in tools/solve_study512.py).  The operator is assembled on the device (z-slab per rank,
    generator of tests/poisson3d-fd/poisson3d_fd.cpp:108-139 on a uniform grid)."""
    import torch
    import blasted_b200 as bb
    from blasted_b200 import solverfactory as sf
    from blasted_b200.dist import Comm, DistMatrix, poisson3d_slab_device

    comm = Comm.from_torch_distributed() if world > 1 else Comm.single()
    t_setup = time.perf_counter()
    part, diag_view = poisson3d_slab_device(n, rank, world)
    A = DistMatrix(comm, part, diag_view)
    s = bb.AsyncSolverSettings(prectype=sf.SOLVER_TYPES["ilu0"], bs=1, nbuildsweeps=FGMRES_SWEEPS[0],
                               napplysweeps=FGMRES_SWEEPS[1])
    prec = bb.SRFactory().create_preconditioner(A.diag, s)
    nloc = A.local_dim()
    ones = torch.ones(nloc, dtype=torch.float64, device="cuda")
    b = A.apply(ones)
    x = torch.zeros_like(b)
    prec.compute()                                         # setup (pattern on device) + warm-up
    # warm-up: one full restart cycle touches every basis vector and every kernel of the solve
    A.solve("fgmres", prec, b, x, tol=1e-8, maxiter=31, restart=30)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    best, info = None, None
    for _ in range(reps):
        x.zero_()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prec.compute()
        info = A.solve("fgmres", prec, b, x, tol=1e-8, maxiter=6000, restart=30)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item()) if best is None else min(best, float(t.item()))
    errt = (x - 1.0).abs().max().reshape(1)
    if dist is not None:
        dist.all_reduce(errt, op=dist.ReduceOp.MAX)
    # Where an iteration goes: three restart cycles with CUDA events around every launch class (the
    # events cost a little; `ms_per_iteration` above is measured without them).  Per rank; the
    # all-reduce class includes the wait for the slowest rank.
    xb = torch.zeros_like(b)
    sf.profile_reset()
    sf.profile_enable(True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    binfo = A.solve("fgmres", prec, b, xb, tol=1e-30, maxiter=90, restart=30)
    e1.record()
    torch.cuda.synchronize()
    sf.profile_enable(False)
    its = max(binfo.iters, 1)
    prof = sf.profile_get_all()
    mine = {k: round(v[0]/its, 4) for k, v in prof.items() if v[1] > 0}
    mine["wall"] = round(e0.elapsed_time(e1)/its, 4)
    mine["unaccounted"] = round(mine["wall"] - sum(v for k, v in mine.items() if k != "wall"), 4)
    sf.profile_reset()
    if dist is not None:
        every = [None]*world
        dist.all_gather_object(every, mine)
    else:
        every = [mine]
    keys = sorted({k for d in every for k in d})
    breakdown = {"iterations": its, "note": "ms per iteration by launch class, CUDA events per launch",
                 "rank0": every[0],
                 "max_over_ranks": {k: max(d.get(k, 0.0) for d in every) for k in keys},
                 "min_over_ranks": {k: min(d.get(k, 0.0) for d in every) for k in keys}}
    mem = torch.cuda.max_memory_allocated()/2**30
    free, total = torch.cuda.mem_get_info()
    return {"problem": f"7-point Poisson {n}^3, z-slabs over {world} GPU(s), block-Jacobi async ILU(0) "
                       f"({FGMRES_SWEEPS[0]},{FGMRES_SWEEPS[1]}) + FGMRES(30), rel. tol 1e-8",
            "scaling": "strong", "sweeps": list(FGMRES_SWEEPS),
            "unknowns": n**3, "iterations": info.iters, "converged": bool(info.converged),
            "time_to_solve_ms": best, "ms_per_iteration": best/max(info.iters, 1),
            "factor_included": True, "max_abs_error": float(errt.item()),
            "timed_solves": reps, "setup_and_warmup_s": t_setup,
            "hbm_used_gib_rank0": (total - free)/2**30, "breakdown": breakdown}


_STDOUT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout: everything libraries print there while the bench runs
    (NCCL's version banner under NCCL_DEBUG=VERSION, for one) is sent to stderr instead."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, default=1024, help="cells per side (C2 = 1024)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1/C3/C4 per-kernel block (N = 1)")
    ap.add_argument("--fgmres-n", type=int, default=512,
                    help="grid size of the FGMRES time-to-solve problem (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
