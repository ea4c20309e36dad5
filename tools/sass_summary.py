"""Per-kernel SASS evidence (run on the CPU box): counts of the mnemonics that show what a kernel
is made of - TMA bulk copies (UBLKCP) and their mbarriers (SYNCS), 256-bit global accesses
(LDG/STG ...256, sm_100 only), fp64 FMAs, shuffles - from `cuobjdump -sass` of the built library.
No tensor-core mnemonics (UTC*MMA, LDTM) are expected: the path is bandwidth-bound by design.

    python tools/sass_summary.py > profiles/sass_summary_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
lib = os.path.join(ROOT, "blasted_b200", "libblasted_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = collections.OrderedDict([
    ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.256", r"\bLDG\.[A-Z0-9.]*256"),
    ("STG.256", r"\bSTG\.[A-Z0-9.]*256"), ("LDG.128", r"\bLDG\.[A-Z0-9.]*128"), ("LDG", r"\bLDG\b"),
    ("STG", r"\bSTG\b"), ("LDS", r"\bLDS\b"), ("DFMA", r"\bDFMA\b"), ("DMUL/DADD", r"\bD(MUL|ADD)\b"),
    ("SHFL", r"\bSHFL\b"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED|REDG)\b"), ("UTC*MMA", r"\bUTC[A-Z]*MMA"),
    ("LDTM", r"\bLDTM\b"), ("HMMA", r"\bHMMA\b")])
arch = re.findall(r"arch = (sm_\w+)", sass)
funcs = re.split(r"\n\s*Function : ", sass)[1:]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("(anonymous namespace)::", "")
    dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("b200::", "")
    if "cub::" in dem or "thrust::" in dem:
        continue
    body = f.split("\n", 1)[1]
    counts = [len(re.findall(p, body)) for p in PAT.values()]
    ninstr = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, flags=re.M))
    rows.append((dem, ninstr, counts))
print(f"# SASS summary of blasted_b200/libblasted_b200.so (cuobjdump -sass; arch {sorted(set(arch))}); CUB kernels omitted")
print("# columns: instructions, then counts of " + ", ".join(PAT.keys()))
tot = [0]*len(PAT)
for dem, n, c in sorted(rows):
    print(f"{dem[:86]:86s} {n:6d} " + " ".join(f"{v:5d}" for v in c))
    tot = [a + b for a, b in zip(tot, c)]
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':86s} {sum(r[1] for r in rows):6d} " + " ".join(f"{v:5d}" for v in tot))
