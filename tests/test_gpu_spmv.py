"""GPU parity: CSR/BSR SpMV and gemv3 through the C ABI vs the oracle and the golden vectors
(deterministic path: 1e-12 relative, BASELINE.json north_star)."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import matgen
from oracle import orc
from util import CASES, case, golden_outputs, golden_matrices, fixture_csr, relerr, SEED

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.mark.parametrize("key", CASES)
def test_spmv_gemv3_fixtures(key):
    g, m = golden_outputs(), case(key)
    A = bb.SRMatrixView(m)
    r = g[key + "_r"]
    assert relerr(A.apply(r), g[key + "_spmv"]) < TOL
    y = np.cos(np.arange(m.dim))
    assert relerr(A.gemv3(0.3, r, -1.2, y), g[key + "_gemv3"]) < TOL
    # z may alias y (Richardson/BiCGSTAB use of gemv3)
    z = y.copy()
    A.gemv3(0.3, r, -1.2, z, z)
    assert relerr(z, g[key + "_gemv3"]) < TOL


@pytest.mark.parametrize("name,bs,rowmajor", [("DK01R", 1, False), ("DK01R", 7, False),
                                              ("DK01R", 7, True), ("small_block3", 3, False),
                                              ("small_block3", 3, True), ("small_block3", 1, False)])
def test_spmv_reference_products(name, bs, rowmajor):
    """b = A x_file to 10 eps-level (tests/mat_ops/testcsrmatrix.cpp:34-36, testbsrmatrix.cpp:46-48)."""
    gm = golden_matrices()
    m = fixture_csr(name, strict=False)
    if bs > 1:
        m = matgen.csr_to_bsr(m, bs, rowmajor, strict_diag=False)
    b = bb.SRMatrixView(m).apply(gm[name + "_x"])
    assert relerr(b, gm[name + "_b"]) < 1e-13


@pytest.mark.parametrize("mk", [lambda: matgen.poisson3d(40), lambda: matgen.poisson3d(20, 27),
                                lambda: matgen.poisson2d(3, 1), lambda: matgen.poisson3d(1),
                                lambda: matgen.block_stencil((64, 48), 4, SEED),
                                lambda: matgen.block_stencil((64, 48), 4, SEED, rowmajor=True),
                                lambda: matgen.block_stencil((12, 11, 10), 5, SEED)])
def test_spmv_synthetic_vs_oracle(mk):
    m = mk()
    rng = np.random.default_rng(SEED)
    x, y = rng.standard_normal(m.dim), rng.standard_normal(m.dim)
    A = bb.SRMatrixView(m)
    assert relerr(A.apply(x), orc().spmv(m, x)) < TOL
    assert relerr(A.gemv3(-1.0, x, 1.0, y), orc().gemv3(m, -1.0, x, 1.0, y)) < TOL


def test_ragged_rows_and_empty_rows():
    """Irregular row lengths incl. empty rows and one dense row."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    a = sp.random(300, 300, density=0.02, random_state=5, format="lil")
    a[7, :] = rng.standard_normal(300)
    a[11, :] = 0
    m = matgen.from_scipy(sp.csr_matrix(a), strict_diag=False)
    x = rng.standard_normal(300)
    assert relerr(bb.SRMatrixView(m).apply(x), orc().spmv(m, x)) < TOL


def test_empty_matrix():
    m = matgen.SRMatrix(0, 1, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0),
                        np.zeros(0, np.int32))
    A = bb.SRMatrixView(m)
    assert A.dim() == 0
    assert A.apply(np.zeros(0)).shape == (0,)


def test_device_pointer_path_and_linearity():
    import torch
    m = matgen.block_stencil((96, 96), 4, SEED)
    A = bb.SRMatrixView(m)
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(m.dim), rng.standard_normal(m.dim)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    ax, ay = A.apply(dx), A.apply(dy)
    axy = A.apply(2.0*dx - 3.0*dy)
    torch.cuda.synchronize()
    assert relerr(ax.cpu().numpy(), orc().spmv(m, x)) < TOL
    assert relerr(axy.cpu().numpy(), (2.0*ax - 3.0*ay).cpu().numpy()) < 1e-12


def test_update_values():
    m = matgen.block_stencil((20, 20), 4, SEED)
    A = bb.SRMatrixView(m)
    x = np.ones(m.dim)
    y1 = A.apply(x)
    A.update_values(2.0*m.vals)
    assert relerr(A.apply(x), 2.0*y1) < 1e-15
