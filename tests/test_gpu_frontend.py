"""GPU parity: the device front end (coordinate -> CSR/BSR, reordering, scaling) through the C ABI
vs the oracle (which tests/test_frontend_cpu.py pins against the reference's COOMatrix /
Reordering / ReorderingScaling).  Integer arrays and moved values: bit-exact."""
import os

import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from blasted_b200.frontend import (COOMatrix, getSRMatrixFromCOO, Reordering, ReorderingScaling,
                                   MatrixReadException, FORWARD, INVERSE, ROW, COLUMN)
from blasted_b200.solverfactory import SOLVER_TYPES
from oracle import orc
from util import fixture_csr, SEED

pytestmark = pytest.mark.gpu


def load(name, bs, rowmajor=False):
    m = fixture_csr(name, strict=False)
    return m if bs == 1 else matgen.csr_to_bsr(m, bs, rowmajor, strict_diag=False)


def triplets(m, seed=0):
    sp = m.to_scipy().tocoo()
    p = np.random.default_rng(seed).permutation(sp.nnz)
    return sp.row[p].astype(np.int32), sp.col[p].astype(np.int32), sp.data[p].astype(np.float64)


def same_matrix(view, browptr, bcolind, diagind, vals):
    h = view.to_host()
    assert np.array_equal(h.browptr, browptr)
    assert np.array_equal(h.bcolind, bcolind)
    assert np.array_equal(h.vals, vals)
    if diagind is not None:
        assert np.array_equal(h.diagind, diagind)


@pytest.mark.parametrize("name,bs,rowmajor", [("2dcyl1", 1, False), ("2dcyl1", 4, False),
                                              ("2dcyl1", 4, True), ("small_block3", 3, False),
                                              ("small_block3", 3, True), ("small_block3", 1, False),
                                              ("msc00726", 1, False), ("DK01R", 7, False),
                                              ("DK01R", 7, True), ("DK01R", 1, False)])
def test_coo_convert_fixtures(name, bs, rowmajor):
    m = load(name, 1)
    r, c, v = triplets(m)
    coo = COOMatrix.from_triplets(m.dim, r, c, v)
    view = getSRMatrixFromCOO(coo, bs, "rowmajor" if rowmajor else "colmajor")
    same_matrix(view, *orc().coo_convert(m.dim, r, c, v, bs, rowmajor))
    # and the result is a working operator
    x = np.cos(np.arange(m.dim))
    assert np.allclose(view.apply(x), m.to_scipy() @ x, rtol=1e-12, atol=1e-12 * np.abs(x).max())


def test_coo_convert_bs5_synthetic_and_device_input():
    import torch
    m = matgen.block_stencil((7, 6, 5), 5, SEED)
    r, c, v = triplets(m, 3)
    view = COOMatrix.from_triplets(m.dim, r, c, v).convertToBSR(5, "colmajor")
    same_matrix(view, m.browptr, m.bcolind, m.diagind, m.vals)
    # triplets already on the device
    import ctypes as C
    from blasted_b200._lib import lib, check
    dr, dc, dv = (torch.as_tensor(a, device="cuda") for a in (r, c, v))
    h = C.c_void_p()
    check(lib.b200_mat_create_coo(m.dim, len(v), C.c_void_p(dr.data_ptr()), C.c_void_p(dc.data_ptr()),
                                  C.c_void_p(dv.data_ptr()), 5, 0, 1, C.byref(h)))
    same_matrix(bb.SRMatrixView.from_handle(h, 5, False), m.browptr, m.bcolind, m.diagind, m.vals)


def test_coo_convert_ragged_blocks_empty_rows_and_errors():
    # ragged: scalar rows of one block row touch different block columns; an empty block row;
    # a row without diagonal.  Block columns come out sorted (the reference appends them in order
    # of first appearance, src/coomatrix.cpp:330-348), same set of blocks, same values.
    n, bs = 12, 3
    r = np.array([0, 0, 1, 2, 2, 9, 10, 11, 11, 3], dtype=np.int32)
    c = np.array([0, 9, 4, 1, 11, 0, 10, 5, 9, 7], dtype=np.int32)
    v = np.arange(1, 11, dtype=np.float64)
    view = COOMatrix.from_triplets(n, r, c, v).convertToBSR(bs, "colmajor")
    h = view.to_host()
    obp, obc, odi, ov = orc().coo_convert(n, r, c, v, bs, False)
    assert np.array_equal(h.browptr, [0, 3, 4, 4, 7])
    for i in range(4):
        assert np.all(np.diff(h.bcolind[h.browptr[i]:h.browptr[i + 1]]) > 0)
        mine = {int(h.bcolind[j]): h.vals[j * 9:(j + 1) * 9] for j in range(h.browptr[i], h.browptr[i + 1])}
        theirs = {int(obc[j]): ov[j * 9:(j + 1) * 9] for j in range(obp[i], obp[i + 1])} if i != 2 else {}
        assert mine.keys() == theirs.keys()
        for k in mine:
            assert np.array_equal(mine[k], theirs[k])
    assert list(h.diagind) == [0, -1, -1, 6]
    # empty matrix
    e = COOMatrix.from_triplets(8, [], [], []).convertToBSR(4)
    assert e.to_host().nnzb == 0 and list(e.to_host().browptr) == [0, 0, 0]
    with pytest.raises(RuntimeError, match="out of range"):
        COOMatrix.from_triplets(4, [0, 4], [0, 1], [1.0, 2.0]).convertToCSR()
    with pytest.raises(RuntimeError, match="multiple of the block size"):
        COOMatrix.from_triplets(10, [0], [0], [1.0]).convertToBSR(4)
    with pytest.raises(RuntimeError, match="invalid storage order"):
        COOMatrix.from_triplets(8, [0], [0], [1.0]).convertToBSR(4, "diagonal")


def test_read_matrix_market(tmp_path):
    m = load("2dcyl1", 1)
    r, c, v = triplets(m, 1)
    path = os.path.join(tmp_path, "a.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% comment\n%\n")
        f.write(f"{m.dim} {m.dim} {len(v)}\n")
        for i in range(len(v)):
            f.write(f"{r[i] + 1} {c[i] + 1} {float(v[i])!r}\n")
    coo = COOMatrix()
    coo.readMatrixMarket(path)
    assert (coo.numrows(), coo.numcols(), coo.numnonzeros()) == (m.dim, m.dim, len(v))
    b4 = matgen.csr_to_bsr(m, 4)
    same_matrix(coo.convertToBSR(4), b4.browptr, b4.bcolind, b4.diagind, b4.vals)
    for banner, msg in [("%%MatrixMarket matrix array real general", "coordinate storage"),
                        ("%%MatrixMarket matrix coordinate pattern general", "pattern"),
                        ("%%MatrixMarket matrix coordinate real symmetric", "general matrices")]:
        bad = os.path.join(tmp_path, "bad.mtx")
        with open(bad, "w") as f:
            f.write(banner + "\n2 2 1\n1 1 1.0\n")
        with pytest.raises(MatrixReadException, match=msg):
            COOMatrix().readMatrixMarket(bad)


@pytest.mark.parametrize("case", [("2dcyl1", 1), ("2dcyl1", 4), ("synth5", 5), ("msc00726", 1)])
@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("which", ["row", "col", "both"])
def test_reorder_matrix_and_vectors(case, inverse, which):
    name, bs = case
    m = matgen.block_stencil((6, 5, 4), 5, SEED) if name == "synth5" else load(name, bs)
    rng = np.random.default_rng(11)
    rord = rng.permutation(m.nbrows).astype(np.int32) if which in ("row", "both") else None
    cord = rng.permutation(m.nbrows).astype(np.int32) if which in ("col", "both") else None
    mode = INVERSE if inverse else FORWARD
    ro = Reordering(bs)
    ro.setOrdering(rord, cord, m.nbrows)
    view = bb.SRMatrixView(m)
    ro.applyOrdering(view, mode)
    bp, bc, v = orc().reorder_matrix(m, rord, cord, inverse)
    same_matrix(view, bp, bc, matgen.find_diagind(bp, bc, strict=False), v)
    # vectors: host arrays and device tensors
    import torch
    vec = rng.standard_normal(m.dim)
    for d, o in ((ROW, rord), (COLUMN, cord)):
        want = vec.copy() if o is None else orc().reorder_vector(bs, o, vec, inverse)
        assert np.array_equal(ro.applyOrdering(vec.copy(), mode, d), want)
        t = torch.as_tensor(vec, device="cuda").clone()
        assert np.array_equal(ro.applyOrdering(t, mode, d).cpu().numpy(), want)


@pytest.mark.parametrize("case", [("2dcyl1", 1), ("2dcyl1", 4), ("synth5", 5)])
@pytest.mark.parametrize("inverse", [False, True])
def test_scaling_matrix_and_vectors(case, inverse):
    name, bs = case
    m = matgen.block_stencil((6, 5, 4), 5, SEED) if name == "synth5" else load(name, bs)
    rng = np.random.default_rng(12)
    rs, cs = rng.uniform(0.5, 2.0, m.nbrows), rng.uniform(0.5, 2.0, m.nbrows)
    mode = INVERSE if inverse else FORWARD
    sc = ReorderingScaling(bs)
    sc.setScaling(rs, cs)
    view = bb.SRMatrixView(m)
    sc.applyScaling(view, mode)
    assert np.array_equal(view.to_host().vals, orc().scale_matrix(m, rs, cs, inverse))
    vec = rng.standard_normal(m.dim)
    assert np.array_equal(sc.applyScaling(vec.copy(), mode, ROW), orc().scale_vector(bs, rs, vec, inverse))
    assert np.array_equal(sc.applyScaling(vec.copy(), mode, COLUMN), orc().scale_vector(bs, cs, vec, inverse))
    # only one of the two set: the other direction is a no-op
    only = ReorderingScaling(bs)
    only.setScaling(rs, None)
    assert np.array_equal(only.applyScaling(vec.copy(), mode, COLUMN), vec)


def test_reordered_solve_round_trip():
    """The use the front end is for: permute (reverse Cuthill-McKee), scale, factor and solve the
    permuted system on the device, map the solution back: same answer as the direct solve."""
    import scipy.sparse.csgraph as cg
    m = load("2dcyl1", 4)
    rng = np.random.default_rng(13)
    perm = cg.reverse_cuthill_mckee(_block_graph(m), symmetric_mode=False).astype(np.int32)
    b = rng.standard_normal(m.dim)
    s = bb.AsyncSolverSettings(prectype=SOLVER_TYPES["seqilu0"], bs=4, nbuildsweeps=1, napplysweeps=1)

    def solve(view, rhs):
        p = bb.SRFactory().create_preconditioner(view, s)
        p.compute()
        x = np.zeros(m.dim)
        solver = bb.GCR(view, p, 30)
        solver.setParams(1e-10, 500)
        info = solver.solve(rhs, x)
        assert info.converged
        return x, info.iters

    x_direct, _ = solve(bb.SRMatrixView(m), b)
    ro = ReorderingScaling(4)
    ro.setOrdering(perm, perm, m.nbrows)
    view = bb.SRMatrixView(m)
    ro.applyOrdering(view, FORWARD)                  # P A P^T
    bp = ro.applyOrdering(b.copy(), FORWARD, ROW)    # P b
    xp, _ = solve(view, bp)
    x_back = ro.applyOrdering(xp, INVERSE, COLUMN)
    assert np.abs(x_back - x_direct).max() < 1e-7 * np.abs(x_direct).max()
    # inverse ordering restores the matrix bit for bit
    ro.applyOrdering(view, INVERSE)
    same_matrix(view, m.browptr, m.bcolind, m.diagind, m.vals)


def _block_graph(m):
    import scipy.sparse as sp
    return sp.csr_matrix((np.ones(m.nnzb), m.bcolind, m.browptr), shape=(m.nbrows, m.nbrows))


def test_full_size_c2_conversion_and_round_trip():
    """BASELINE C2 size (1024 x 1024 cells, bs = 4; 83.8 M scalar triplets): the device conversion
    reproduces the generator's BSR arrays bit for bit; a random symmetric permutation followed by
    its inverse restores them (size-independent property)."""
    import torch
    m = matgen.block_stencil((1024, 1024), 4, 1)
    nb = m.nbrows
    # scalar triplets of the block matrix, built on the device in a scrambled order
    rows = torch.repeat_interleave(torch.arange(nb, device="cuda", dtype=torch.int32),
                                   torch.as_tensor(np.diff(m.browptr), device="cuda"))
    cols = torch.as_tensor(m.bcolind, device="cuda")
    k = torch.arange(16, device="cuda", dtype=torch.int32)
    # column-major blocks: entry k of a block is (row k % 4, column k // 4)
    r = (rows[:, None] * 4 + (k % 4)[None, :]).reshape(-1)
    c = (cols[:, None] * 4 + (k // 4)[None, :]).reshape(-1)
    v = torch.as_tensor(m.vals, device="cuda")
    p = torch.randperm(r.numel(), device="cuda")
    r, c, v = r[p].contiguous(), c[p].contiguous(), v[p].contiguous()
    del p
    import ctypes as C
    from blasted_b200._lib import lib, check
    h = C.c_void_p()
    check(lib.b200_mat_create_coo(m.dim, r.numel(), C.c_void_p(r.data_ptr()), C.c_void_p(c.data_ptr()),
                                  C.c_void_p(v.data_ptr()), 4, 0, 1, C.byref(h)))
    del r, c, v
    view = bb.SRMatrixView.from_handle(h, 4, False)
    same_matrix(view, m.browptr, m.bcolind, m.diagind, m.vals)
    perm = np.random.default_rng(2).permutation(nb).astype(np.int32)
    ro = Reordering(4)
    ro.setOrdering(perm, perm, nb)
    ro.applyOrdering(view, FORWARD)
    hm = view.to_host()
    assert np.array_equal(np.diff(hm.browptr), np.diff(m.browptr)[perm])
    assert np.array_equal(hm.diagind >= 0, np.ones(nb, dtype=bool))
    ro.applyOrdering(view, INVERSE)
    same_matrix(view, m.browptr, m.bcolind, m.diagind, m.vals)


@pytest.mark.parametrize("case", [("2dcyl1", 1), ("2dcyl1", 4), ("DK01R", 7)])
@pytest.mark.parametrize("inverse", [False, True])
def test_reordering_scaling_adapter_behind_reference_interface(case, inverse):
    """B200ReorderingScaling<bs> (blasted_b200/host) derives from the reference's
    ReorderingScaling<double,int,bs>; driven through the reference's own virtuals it must give the
    reference's results bit for bit (matrix arrays, row- and column-direction vectors)."""
    import oracle
    if not oracle.have_ref():
        pytest.skip("reference build not present")
    name, bs = case
    m = load(name, bs)
    rng = np.random.default_rng(21)
    rord = rng.permutation(m.nbrows).astype(np.int32)
    cord = rng.permutation(m.nbrows).astype(np.int32)
    rs, cs = rng.uniform(0.5, 2.0, m.nbrows), rng.uniform(0.5, 2.0, m.nbrows)
    vec = rng.standard_normal(m.dim)
    kw = dict(rord=rord, cord=cord, rowscale=rs, colscale=cs, inverse=inverse, rowvec=vec, colvec=vec)
    want = oracle.ref().reorder_scale(m, **kw)
    got = oracle.ref().reorder_scale(m, through_b200=True, **kw)
    for a, b in zip(want, got):
        assert np.array_equal(a, b)
