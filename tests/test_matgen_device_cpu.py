"""CPU: the torch generators of blasted_b200/matgen_device.py (run on the CPU device here) assemble
the same operators as the numpy generators the parity tests use."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    # by path: importing the package would load the CUDA library, which these tests do not need
    spec = importlib.util.spec_from_file_location("_b200_" + name, os.path.join(ROOT, "blasted_b200", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_b200_" + name] = mod
    spec.loader.exec_module(mod)
    return mod


matgen, md = _load("matgen"), _load("matgen_device")


@pytest.mark.parametrize("dims,stencil,chunk", [((7, 5, 9), 7, 100), ((16, 12, 10), 7, 333),
                                                ((5, 5, 3), 27, 8), ((9, 7, 6), 27, 1 << 21)])
def test_poisson_device_equals_host(dims, stencil, chunk):
    m = matgen.poisson3d(0, stencil, dims=dims)
    bp, bc, slot, offs = md.stencil_pattern_device(dims, stencil == 27, device="cpu", rows_per_chunk=chunk)
    assert np.array_equal(bp.numpy(), m.browptr) and np.array_equal(bc.numpy(), m.bcolind)
    bp, bc, v = md.poisson3d_device(dims, stencil, device="cpu")
    assert np.array_equal(v.numpy(), m.vals)


@pytest.mark.parametrize("dims,bs", [((12, 9), 4), ((6, 5, 4), 5)])
def test_block_stencil_device_follows_the_host_recipe(dims, bs):
    m = matgen.block_stencil(dims, bs, 1)
    nb, bp, bc, v = md.block_stencil_device(dims, bs, 1, device="cpu")
    assert nb == m.nbrows and np.array_equal(bp.numpy(), m.browptr) and np.array_equal(bc.numpy(), m.bcolind)
    blocks = v.numpy().reshape(-1, bs, bs).transpose(0, 2, 1)             # logical [r, c]
    rows = np.repeat(np.arange(nb), np.diff(m.browptr))
    off = np.abs(blocks).sum(axis=2).max(axis=1)
    off[m.diagind] = 0
    assert off.max() <= 0.125*bs + 1e-12
    rowsum = np.bincount(rows, weights=off, minlength=nb)
    d = blocks[m.diagind]
    dg = d[:, np.arange(bs), np.arange(bs)]
    assert np.all(np.abs(dg - (rowsum + 1.0)[:, None]) <= 0.25 + 1e-12)
