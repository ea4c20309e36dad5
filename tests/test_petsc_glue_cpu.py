"""CPU: the PETSc-typed half of the PCSHELL glue (examples/petsc_glue/blasted_b200_petsc.c) compiles
as strict C99 against a declaration-only PETSc header with PETSc's published signatures, defines
every C symbol of the reference's include/blasted_petsc.h:88-167 and links against
libblasted_b200.so with only PETSc symbols left undefined.  (PETSc itself is not in the image.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "examples", "petsc_glue", "blasted_b200_petsc.c")
STUB = os.path.join(ROOT, "examples", "petsc_glue", "petsc_stub")

# the extern "C" block of include/blasted_petsc.h
REFERENCE_SYMBOLS = ["newBlastedDataList", "computeTotalTimes", "destroyBlastedDataList",
                     "setup_blasted_stack", "newBlastedDataContext", "appendBlastedDataContext",
                     "setup_localpreconditioner_blasted", "cleanup_blasted",
                     "compute_preconditioner_blasted", "apply_local_blasted", "relax_local_blasted"]


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_petsc_glue_compiles_and_defines_the_reference_symbols(tmp_path):
    obj = tmp_path / "glue.o"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fPIC",
                    f"-I{STUB}", f"-I{os.path.join(ROOT, 'include')}", "-c", SRC, "-o", str(obj)], check=True)
    nm = subprocess.run(["nm", str(obj)], check=True, capture_output=True, text=True).stdout
    defined = {l.split()[-1] for l in nm.splitlines() if " T " in l}
    undefined = {l.split()[-1] for l in nm.splitlines() if " U " in l}
    assert set(REFERENCE_SYMBOLS) <= defined
    # everything it needs besides libc is either the device library's shell API or PETSc
    lib = os.path.join(ROOT, "blasted_b200", "libblasted_b200.so")
    exported = subprocess.run(["nm", "-D", "--defined-only", lib], check=True, capture_output=True,
                              text=True).stdout
    exported = {l.split()[-1] for l in exported.splitlines()}
    ours = {s for s in undefined if s.startswith("b200_")}
    assert ours and ours <= exported, ours - exported
    petsc = {s for s in undefined if s.startswith(("Petsc", "KSP", "PC", "Mat", "Vec"))}
    libc = undefined - ours - petsc
    assert libc <= {"abort", "fflush", "fprintf", "printf", "puts", "snprintf", "stdout", "stderr", "strcpy",
                    "strncmp", "strstr", "memset", "__stack_chk_fail", "_GLOBAL_OFFSET_TABLE_"}, libc
    # links into a shared object with only the PETSc symbols unresolved
    so = tmp_path / "libblasted_b200_petsc.so"
    subprocess.run(["gcc", "-shared", "-o", str(so), str(obj), f"-L{os.path.dirname(lib)}", "-lblasted_b200",
                    "-Wl,--unresolved-symbols=ignore-all"], check=True)
    assert so.exists()
