"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "blasted_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from blasted_b200 import _lib
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/blasted_b200.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(_lib.SYMBOLS) == names
    # the PETSc-free PCSHELL core (include/blasted_b200_shell.h) lives in the same library
    txt = open(os.path.join(ROOT, "include", "blasted_b200_shell.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    shell = sorted(set(re.findall(r"\b(b200_shell_[a-z0-9_]+)\s*\(", txt)))
    assert len(shell) >= 15
    for n in shell:
        assert hasattr(_lib.lib, n), f"{n} declared in include/blasted_b200_shell.h but not exported"


def test_no_oracle_in_product_path():
    """The product never links or imports the oracle."""
    pkg = os.path.join(ROOT, "blasted_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "blasted_oracle" not in txt
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M)


def test_compute_fails_loudly_without_gpu():
    import blasted_b200 as bb
    if bb.device_count() > 0:
        pytest.skip("GPU present")
    m = bb.matgen.poisson3d(3)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        bb.CSRMatrixView(m)


def test_settings_struct_layout():
    from blasted_b200 import _lib
    assert ctypes.sizeof(_lib.Settings) == 12*4
    assert ctypes.sizeof(_lib.SolveInfo) == 2*4 + 4*8


def test_factory_strings():
    import blasted_b200 as bb
    f = bb.SRFactory()
    assert f.solverTypeFromString("ilu0") == 3 and f.solverTypeFromString("none") == 10
    with pytest.raises(ValueError):
        f.solverTypeFromString("bogus")


def _build_c_example(tmp_path):
    """The headers are plain C99 and the library links from C (what cgo / JNI / ctypes stubs need)."""
    import subprocess
    exe = os.path.join(tmp_path, "c_abi_example")
    cmd = ["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror",
           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_example.c"),
           "-L" + os.path.join(ROOT, "blasted_b200"), "-lblasted_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "blasted_b200"), "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_c_example_compiles_and_links(tmp_path):
    import subprocess
    import blasted_b200 as bb
    exe = _build_c_example(tmp_path)
    rc = subprocess.run([exe], capture_output=True, text=True).returncode
    assert rc == (0 if bb.device_count() > 0 else 77)


@pytest.mark.gpu
def test_c_example_runs_on_gpu(tmp_path):
    import subprocess
    out = subprocess.run([_build_c_example(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "|z_abi - z_shell| = 0.000e+00" in out.stdout
