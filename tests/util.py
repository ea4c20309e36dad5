"""Shared helpers for the test-suite: golden fixtures and the matrix cases used for parity."""
import functools
import os

import numpy as np

from blasted_b200 import matgen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 20261018


@functools.lru_cache(maxsize=None)
def golden_matrices():
    return dict(np.load(os.path.join(GOLDEN, "matrices.npz")))


@functools.lru_cache(maxsize=None)
def golden_outputs():
    return dict(np.load(os.path.join(GOLDEN, "reference_outputs.npz")))


def fixture_csr(name, strict=True):
    g = golden_matrices()
    m = matgen.SRMatrix(len(g[name + "_browptr"]) - 1, 1, g[name + "_browptr"], g[name + "_bcolind"],
                        g[name + "_vals"], None)
    m.diagind = matgen.find_diagind(m.browptr, m.bcolind, strict)
    return m


@functools.lru_cache(maxsize=None)
def case(key):
    """The matrix cases of tests/golden/make_golden.py, by key."""
    if key == "2dcyl1_csr":
        return fixture_csr("2dcyl1")
    if key == "2dcyl1_bsr4":
        return matgen.csr_to_bsr(fixture_csr("2dcyl1"), 4, False)
    if key == "2dcyl1_bsr4r":
        return matgen.csr_to_bsr(fixture_csr("2dcyl1"), 4, True)
    if key == "msc00726_csr":
        return fixture_csr("msc00726")
    if key == "synth_bsr5":
        return matgen.block_stencil((6, 5, 4), 5, SEED)
    raise KeyError(key)


CASES = ["2dcyl1_csr", "2dcyl1_bsr4", "2dcyl1_bsr4r", "msc00726_csr", "synth_bsr5"]


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
