"""GPU tests of the PETSc-free PCSHELL core (include/blasted_b200_shell.h): setup / apply / relax
through the shell give what the preconditioner objects give directly, values are re-read at every
setup (PETSc rewrites `a` in place), PrecInfo records and times accumulate."""
import numpy as np
import pytest

import blasted_b200 as bb
from blasted_b200 import shell
from blasted_b200.solverfactory import SOLVER_TYPES, FACT_INIT, APPLY_INIT
from oracle import orc
from util import case, relerr

pytestmark = pytest.mark.gpu


def direct(m, name, **kw):
    p = bb.SRFactory().create_preconditioner(bb.SRMatrixView(m), bb.AsyncSolverSettings(
        prectype=SOLVER_TYPES[name], bs=m.bs, **kw))
    p.compute()
    return p


@pytest.mark.parametrize("key", ["2dcyl1_csr", "2dcyl1_bsr4", "synth_bsr5"])
def test_shell_setup_apply_matches_objects(key):
    m = case(key)
    r = np.cos(np.arange(m.dim))
    lst = shell.BlastedDataList()
    # exact variants are deterministic: the shell must reproduce the objects bit for bit
    for pc, sweeps, name, kw in [("seqilu0", (1, 1), "seqilu0", dict(nbuildsweeps=1, napplysweeps=1)),
                                 ("ilu0", (-1, -1), "seqilu0", dict(nbuildsweeps=1, napplysweeps=1)),
                                 ("jacobi", (1, 1), "jacobi", {}),
                                 ("level_sgs", (1, 1), "level_sgs", {})]:
        node = lst.append_new()
        node.set_options(shell.make_options(pc, sweeps, fact_init="init_original", apply_init="init_zero",
                                            precinfo=True))
        node.setup(m)
        z = node.apply(r)
        assert np.array_equal(z, direct(m, name, **kw).apply(r)), (pc, sweeps)
        assert node.node.bs == m.bs and node.node.factorwalltime > 0 and node.node.applywalltime > 0
        assert len(node.infos()) == 1
    # asynchronous: converged sweeps agree with the exact operation
    node = lst.append_new()
    node.set_options(shell.make_options("ilu0", (30, 60), scale=False, fact_init="init_original",
                                        apply_init="init_jacobi"))
    node.setup(m)
    assert relerr(node.apply(r), direct(m, "seqilu0", nbuildsweeps=1, napplysweeps=1).apply(r)) < 1e-8
    assert not node.offers_relaxation()
    fw, aw, _, _ = lst.compute_total_times()
    assert fw > 0 and aw > 0 and lst.c.size == 5
    lst.destroy()


def test_shell_rereads_values_and_relaxes():
    import copy
    m = case("2dcyl1_bsr4")
    lst = shell.BlastedDataList()
    node = lst.append_new()
    # level_sgs: the exact (deterministic) SGS, so that every comparison below can be sharp
    node.set_options(shell.make_options("level_sgs", (1, 1), apply_init="init_zero", precinfo=True))
    node.setup(m)
    r = np.sin(np.arange(m.dim))
    z1 = node.apply(r)
    # new Jacobian values on the same pattern, rewritten in place
    m2 = copy.deepcopy(m)
    m2.vals *= 1.5
    node.setup(m2)
    assert relerr(node.apply(r), z1/1.5) < 1e-12
    assert len(node.infos()) == 2
    # Richardson callback: `it` relaxation steps from a zero guess = the object's apply_relax
    assert node.offers_relaxation()
    x = np.full(m.dim, 7.0)
    its, reason = node.relax(r, x, 3, guesszero=True)
    assert (its, reason) == (3, 4)
    p = direct(m2, "level_sgs")
    p.setApplyParams(maxits=3)
    x_ref = np.zeros(m.dim)
    p.apply_relax(r, x_ref)
    assert np.array_equal(x, x_ref)
    d = orc().jacobi_setup(m2)
    assert relerr(x, orc().sgs_relax(m2, d, 3, r, np.zeros(m.dim))) < 1e-12
    with pytest.raises(RuntimeError, match="changed size"):
        node.setup(case("2dcyl1_csr"))
    node.cleanup()
    with pytest.raises(RuntimeError, match="apply before setup"):
        node.apply(r)
    lst.destroy()


def test_shell_device_pointers_and_errors():
    import torch
    m = case("2dcyl1_bsr4")
    lst = shell.BlastedDataList()
    node = lst.append_new()
    node.set_options(shell.make_options("seqilu0", (1, 1), apply_init="init_zero"))
    node.setup(m)
    r = np.cos(np.arange(m.dim))
    zd = node.apply(torch.as_tensor(r, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(zd.cpu().numpy(), node.apply(r))
    bad = lst.append_new()
    bad.set_options(shell.make_options("ilu0", (1, 1)))
    from blasted_b200 import matgen
    m3 = matgen.block_stencil((4, 4), 3, 1)          # the glue lets bs = 3 through, the factory does not
    with pytest.raises(RuntimeError, match="not supported for column major"):
        bad.setup(m3)
    assert not bad.node.bprec and not bad.node.bmat    # nothing left behind
    m2 = matgen.block_stencil((4, 4), 2, 1)
    with pytest.raises(RuntimeError, match="Block size 2 is not supported"):
        bad.setup(m2)
    lst.destroy()
