"""CPU tests of the PETSc-free PCSHELL core: context list, options -> settings (including the
sequential symbol of src/blasted_petsc.cpp:94-134) and error behaviour.  No GPU needed."""
import pytest

from blasted_b200 import shell
from blasted_b200.solverfactory import SOLVER_TYPES, FACT_INIT, APPLY_INIT


def test_list_management_and_times():
    lst = shell.BlastedDataList()
    a = lst.append_new()
    b = lst.append_new()                       # new nodes go to the head of the list
    assert lst.c.size == 2
    heads = [n.p.contents for n in lst.nodes()]
    assert len(heads) == 2 and not heads[0].first_setup_done
    a.node.factorwalltime, a.node.applywalltime = 1.5, 0.25
    b.node.factorwalltime, b.node.applywalltime, b.node.applycputime = 0.5, 0.75, 2.0
    fw, aw, fc, ac = lst.compute_total_times()
    assert (fw, aw, fc, ac) == (2.0, 1.0, 0.0, 2.0)
    lst.destroy()
    assert lst.c.size == 0


@pytest.mark.parametrize("pc,sweeps,want_type,want_sweeps", [
    ("ilu0", (3, 4), "ilu0", (3, 4)),
    ("ilu0", (-1, 4), "sfilu0", (1, 4)),          # sequential factorisation requested
    ("ilu0", (3, -1), "sapilu0", (3, 1)),         # sequential application requested
    ("ilu0", (-1, -1), "seqilu0", (1, 1)),
    ("sfilu0", (2, -1), "seqilu0", (1, 1)),
    ("sapilu0", (-1, 2), "seqilu0", (1, 1)),
    ("seqilu0", (7, 9), "seqilu0", (7, 9)),
    ("sgs", (1, 3), "sgs", (1, 3)),
    ("jacobi", (5, 5), "jacobi", (1, 1)),         # sweeps are not read for jacobi / level_sgs / none
    ("level_sgs", (5, 5), "level_sgs", (1, 1)),
    ("async_level_ilu0", (2, 1), "async_level_ilu0", (2, 1)),
])
def test_options_to_settings(pc, sweeps, want_type, want_sweeps):
    lst = shell.BlastedDataList()
    node = lst.append_new()
    node.set_options(shell.make_options(pc, sweeps, scale=True, fact_init="init_sgs",
                                        apply_init="init_zero", chunk=64, precinfo=True))
    node.node.bs = 4
    s = node.settings()
    assert s.prectype == SOLVER_TYPES[want_type]
    assert (s.nbuildsweeps, s.napplysweeps) == want_sweeps
    assert s.bs == 4 and s.blockstorage == 0 and s.compute_precinfo == 1 and s.relax == 0
    uses_sweeps = pc not in ("jacobi", "level_sgs", "none")
    fact = pc in ("ilu0", "sapilu0", "async_level_ilu0")
    assert s.scale == (1 if fact else 0)
    # the factor initialisation is looked up for the FINAL type (after the sequential remapping)
    final_fact = want_type in ("ilu0", "sapilu0", "async_level_ilu0")
    if uses_sweeps and final_fact and fact:
        assert s.fact_inittype == FACT_INIT["init_sgs"]
    if uses_sweeps:
        assert s.apply_inittype == APPLY_INIT["init_zero"] and s.thread_chunk_size == 64
    assert node.offers_relaxation() == (pc not in ("ilu0", "cscbgs", "none"))
    lst.destroy()


def test_option_errors():
    lst = shell.BlastedDataList()
    node = lst.append_new()
    with pytest.raises(RuntimeError, match="Preconditioner type not available"):
        node.set_options(shell.make_options("ilu7"))
    node.set_options(shell.make_options("sgs", (-1, 2)))
    with pytest.raises(RuntimeError, match="Seq. fact. only supported"):
        node.settings()
    node.set_options(shell.make_options("sgs", (2, -1)))
    with pytest.raises(RuntimeError, match="Seq. appl. only supported"):
        node.settings()
    node.set_options(shell.make_options("ilu0", (1, 1), fact_init="init_foo"))
    with pytest.raises(RuntimeError, match="Factor initialization not recongnized"):
        node.settings()
    node.set_options(shell.make_options("sgs", (1, 1), apply_init="init_bar"))
    with pytest.raises(RuntimeError, match="Apply initialization not recongnized"):
        node.settings()
    fresh = lst.append_new()
    from blasted_b200 import matgen
    with pytest.raises(RuntimeError, match="set_options must come before"):
        fresh.setup(matgen.poisson3d(3))
    lst.destroy()
