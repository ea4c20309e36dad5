"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref/libblasted_ref.so):

    python tests/golden/make_golden.py

Writes
  matrices.npz          the reference's own small test matrices (tests/input/*, tests/mat_ops/input/*)
                        converted to CSR arrays, with their x / b vectors;
  reference_outputs.npz outputs of the reference objects on those matrices, single-threaded
                        (OMP threads = 1, i.e. the deterministic "sequential" reference results):
                        ILU position lists, level schedules, exact ILU(0) factors, preconditioner
                        applications, SpMV products and Krylov iteration counts.
The GPU box has no /root/reference: tests read only these two files.
"""
import os
import sys
import warnings

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")

from blasted_b200 import matgen          # noqa: E402
from oracle import ref                   # noqa: E402

REF = "/root/reference/tests/"
SEED = 20261018


def vec(path):
    return np.asarray(sio.mmread(path), dtype=np.float64).ravel()


def main():
    R = ref()
    R.set_num_threads(1)
    mats = {
        "2dcyl1": (REF + "input/fvens-2dcyl1/2dcyl1", 4),
        "msc00726": (REF + "input/boeing-msc00726/msc00726", 1),
        "DK01R": (REF + "mat_ops/input/fluorem-dk01r/DK01R", 7),
        "small_block3": (REF + "mat_ops/input/small_block3_matrix", 3),
    }
    mz, out = {}, {}
    for name, (path, bs) in mats.items():
        m = matgen.from_scipy(sio.mmread(path + ".mtx"), strict_diag=False)
        mz[name + "_browptr"] = m.browptr
        mz[name + "_bcolind"] = m.bcolind
        mz[name + "_vals"] = m.vals
        mz[name + "_x"] = vec(path + "_x.mtx")
        mz[name + "_b"] = vec(path + "_b.mtx")
        mz[name + "_bs"] = np.int32(bs)
    np.savez_compressed(os.path.join(HERE, "matrices.npz"), **mz)

    rng = np.random.default_rng(SEED)
    cases = []
    for name in ("2dcyl1", "msc00726"):
        m1 = matgen.SRMatrix(len(mz[name + "_browptr"]) - 1, 1, mz[name + "_browptr"],
                             mz[name + "_bcolind"], mz[name + "_vals"], None)
        m1.diagind = matgen.find_diagind(m1.browptr, m1.bcolind)
        cases.append((name + "_csr", m1))
        if name == "2dcyl1":
            cases.append((name + "_bsr4", matgen.csr_to_bsr(m1, 4, False)))
            cases.append((name + "_bsr4r", matgen.csr_to_bsr(m1, 4, True)))
    # a small bs=5 case (no bs=5 fixture exists in the reference; synthetic, seeded)
    cases.append(("synth_bsr5", matgen.block_stencil((6, 5, 4), 5, SEED)))

    for key, m in cases:
        r = rng.standard_normal(m.dim)
        out[key + "_r"] = r
        # SpMV of the fixture x (or r)
        out[key + "_spmv"] = R.spmv(m, r)
        out[key + "_gemv3"] = R.gemv3(m, 0.3, r, -1.2, np.cos(np.arange(m.dim)))
        posptr, lowerp, upperp = R.ilu_positions(m)
        out[key + "_posptr"], out[key + "_lowerp"], out[key + "_upperp"] = posptr, lowerp, upperp
        out[key + "_levels"] = R.compute_levels(m)
        for scale in (False, True):
            tag = key + ("_scaled" if scale else "")
            p = R.prec(m, "seqilu0", scale=scale, nbuildsweeps=1, napplysweeps=1,
                       fact_init="init_original", apply_init="init_jacobi", compute_precinfo=True)
            info = p.compute()
            out[tag + "_exact_ilu"] = p.factor()          # diagonal blocks inverted for bs>1
            out[tag + "_ilu_apply"] = p.apply(r)
            out[tag + "_precinfo"] = info
            p.close()
        p = R.prec(m, "sgs", napplysweeps=1)
        p.compute()
        out[key + "_dblocks"] = p.dblocks()
        out[key + "_sgs_apply"] = p.apply(r)
        out[key + "_sgs_relax3"] = p.apply_relax(r, np.zeros(m.dim), 3)
        p.close()
        p = R.prec(m, "jacobi")
        p.compute()
        out[key + "_jacobi_apply"] = p.apply(r)
        p.close()

    # Krylov iteration counts of the reference drivers (tests/solvers.cpp), sequential
    # preconditioners, rel. tol 1e-10
    name = "2dcyl1"
    b = mz[name + "_b"]
    its = {}
    for key, m in cases:
        if not key.startswith("2dcyl1") and not key.startswith("msc"):
            continue
        bb = mz[key.split("_")[0] + "_b"]
        for prec in ("seqilu0", "sgs", "jacobi"):
            for solver in ("bicgstab", "gcr"):
                p = R.prec(m, prec, nbuildsweeps=1, napplysweeps=1)
                p.compute()
                x, n, rr, _ = R.solve(solver, p, m, bb, tol=1e-10, maxiter=2000, restart=30)
                its[f"{key}_{prec}_{solver}"] = n
                out[f"its_{key}_{prec}_{solver}"] = np.array([n, rr])
                p.close()
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    for k, v in its.items():
        print(k, v)


if __name__ == "__main__":
    main()
