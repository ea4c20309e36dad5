"""GPU: the drop-in boundary in C++.  The product's B200Factory / B200Preconditioner adapters
(blasted_b200/host) derive from the reference's own FactoryBase / SRPreconditioner; here the
REFERENCE's Krylov drivers (tests/solvers.cpp, compiled unmodified into oracle/_ref) run with the
device preconditioner plugged in, next to the reference's own preconditioner."""
import numpy as np
import pytest

from oracle import have_ref, ref
from util import case, golden_matrices, golden_outputs, relerr

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs the reference tree)")]


@pytest.mark.parametrize("key,prec,kw", [
    ("2dcyl1_bsr4", "seqilu0", {}),
    ("2dcyl1_bsr4", "ilu0", dict(nbuildsweeps=10, napplysweeps=30)),
    ("2dcyl1_bsr4r", "seqilu0", {}),
    ("2dcyl1_csr", "seqilu0", {}),
    ("2dcyl1_bsr4", "jacobi", {}),
    ("msc00726_csr", "seqilu0", {}),
])
def test_reference_bicgstab_with_device_preconditioner(key, prec, kw):
    R, gm, g, m = ref(), golden_matrices(), golden_outputs(), case(key)
    R.set_num_threads(1)
    b = gm[key.split("_")[0] + "_b"]
    pd = R.prec_b200(m, prec, **kw)
    info = pd.compute()
    assert info.shape == (6,)
    # apply parity against the reference object of the exact type
    exact = "seqilu0" if prec == "ilu0" else prec
    pr = R.prec(m, exact, nbuildsweeps=1, napplysweeps=1)
    pr.compute()
    r = g[key + "_r"]
    assert relerr(pd.apply(r), pr.apply(r)) < (1e-9 if prec == "ilu0" else 1e-12)
    assert pd.dim() == pr.dim() == m.dim
    # the reference's own BiCGSTAB on both
    xd, itd, rrd, _ = R.solve("bicgstab", pd, m, b, tol=1e-10, maxiter=2000)
    xr, itr, rrr, _ = R.solve("bicgstab", pr, m, b, tol=1e-10, maxiter=2000)
    assert rrd < 1e-10
    tol = 0.15 if (key.startswith("msc") and prec == "jacobi") else 0.05
    assert abs(itd - itr) <= max(1, int(np.ceil(tol*itr))), (itd, itr)


def test_adapter_error_types():
    R, m = ref(), case("2dcyl1_bsr4")
    p = R.prec_b200(m, "ilu0")
    p.compute()
    with pytest.raises(RuntimeError, match="ILU relaxation not implemented!"):
        p.apply_relax(np.ones(m.dim), np.zeros(m.dim), 2)
    with pytest.raises(ValueError):
        R.prec_b200(m, "bogus")
