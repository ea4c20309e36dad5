"""Turn the ncu reports in gpurun_out/ into the committed summaries under profiles/ (run on the CPU box)."""
import csv, json, os, subprocess, sys, collections
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size"]
CLASS = {"block_ilu0_lower": "factor_lower", "block_ilu0_upper": "factor_upper",
         "block5_lower_staged": "factor_lower", "block5_upper_staged": "factor_upper",
         "tri5_staged_kernel<0": "tri_lower", "tri5_staged_kernel<1": "tri_upper",
         "tri_block_kernel<4, 0": "tri_lower", "tri_block_kernel<4, 1": "tri_upper",
         "tri_block_pipe_kernel<4, 0": "tri_lower", "tri_block_pipe_kernel<4, 1": "tri_upper",
         "tri_block_kernel<5, 0": "tri_lower", "tri_block_kernel<5, 1": "tri_upper",
         "scalar_lower_kernel": "factor_lower", "scalar_upper_kernel": "factor_upper",
         "csr_stream_kernel<2,": "tri_lower", "csr_stream_kernel<3,": "tri_upper",
         "csr_stream_kernel<0,": "spmv", "bsr": "spmv", "csr_spmv": "spmv"}
traffic = {}
algbytes = {}
for w in ("c1", "c2", "c3", "c4", "c3s", "p128"):
    ab = os.path.join(ROOT, "gpurun_out", f"algbytes_{w}_{R}.txt")
    if os.path.exists(ab):
        for tok in open(ab).read().split():
            if "=" in tok:
                k, v = tok.split("=")
                algbytes.setdefault(w, {})[k] = int(v)
    rawf = os.path.join(ROOT, "gpurun_out", f"raw_{w}_{R}.csv")
    if not os.path.exists(rawf):
        continue
    rows = list(csv.reader(open(rawf).read().splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"ncu_{w}_{R}.csv"), "w") as f:
        wr = csv.writer(f)
        cols = ["Kernel Name"] + [c for c in WANT if c in hdr]
        stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
        wr.writerow(cols + ["top_stalls"])
        wr.writerow([""] + [units[hdr.index(c)] for c in cols[1:]] + [""])
        for r in rows[2:]:
            st = sorted(((float(r[hdr.index(h)].replace(",", "") or 0), h[34:-24]) for h in stall), reverse=True)[:3]
            wr.writerow([r[hdr.index("Kernel Name")][:90]] + [r[hdr.index(c)] for c in cols[1:]] +
                        ["; ".join(f"{n}={v:.1f}" for v, n in st)])
            if True:
                name = r[hdr.index("Kernel Name")]
                for key, cls in CLASS.items():
                    if key in name:
                        def gb(x, u):
                            v = float(x.replace(",", ""))
                            return v*{"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
                        rd = gb(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
                        wrb = gb(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
                        traffic.setdefault(w.upper(), {}).setdefault(cls, []).append(rd + wrb)
                        break
if traffic:
    out = {}
    for w, d in traffic.items():
        out[w] = {}
        for cls, v in d.items():
            t = sum(v)/len(v)
            a = algbytes.get(w.lower(), {}).get(cls)
            out[w][cls] = {"dram_bytes_per_launch": t, "algorithmic_bytes_per_launch": a,
                           "dram_over_algorithmic": (t/a if a else None)}
    json.dump(out, open(os.path.join(OUT, f"traffic_{R}.json"), "w"), indent=1)
# launch list: per-kernel share of the bench step
ll = os.path.join(ROOT, "gpurun_out", f"launches_{R}.csv")
if os.path.exists(ll):
    lines = [l for l in open(ll) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi: continue
        t = float(r[vi].replace(",", ""))*{"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(r[ui], 1.0)
        k = r[ki][:80]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"launches_{R}_summary.csv"), "w") as f:
        wr = csv.writer(f)
        wr.writerow(["kernel", "launches", "total_us", "avg_us", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            wr.writerow([k, a[0], f"{a[1]:.1f}", f"{a[1]/a[0]:.2f}", f"{a[1]/tot:.3f}"])
    import shutil
    shutil.copy(ll, os.path.join(OUT, f"launches_{R}.csv"))
print("profiles written to", OUT, os.listdir(OUT))
