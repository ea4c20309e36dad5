"""PETSc-free core of BLASTed's PCSHELL glue (include/blasted_b200_shell.h) from Python.

Mirrors what src/blasted_petsc.cpp does between PETSc's callbacks and the preconditioner object:
`BlastedData.set_options` = setupDataFromOptions (:137-208), `setup` = compute_preconditioner_blasted
(:403-429, creating the object on the first call as createNewPreconditioner :216-311 does),
`apply` = apply_local_blasted (:474-517), `relax` = relax_local_blasted (:519-576), and
`BlastedDataList` = Blasted_data_list with computeTotalTimes (:723-735).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib, check, Settings

OPT_STRLEN = 20
SEQUENTIAL_SYMBOL = -1


class ShellOptions(C.Structure):
    _fields_ = [("pc_type", C.c_char * OPT_STRLEN), ("async_sweeps", C.c_int * 2),
                ("use_symmetric_scaling", C.c_int), ("fact_init_type", C.c_char * OPT_STRLEN),
                ("apply_init_type", C.c_char * OPT_STRLEN), ("thread_chunk_size", C.c_int),
                ("compute_preconditioner_info", C.c_int)]


class ShellNode(C.Structure):
    pass


ShellNode._fields_ = [("bprec", C.c_void_p), ("bmat", C.c_void_p), ("bs", C.c_int),
                      ("prectypestr", C.c_char * OPT_STRLEN), ("prectype", C.c_int), ("scale", C.c_int),
                      ("threadchunksize", C.c_int), ("nbuildsweeps", C.c_int), ("napplysweeps", C.c_int),
                      ("factinittype", C.c_char * OPT_STRLEN), ("applyinittype", C.c_char * OPT_STRLEN),
                      ("compute_precinfo", C.c_int), ("infolist", C.c_void_p),
                      ("first_setup_done", C.c_int), ("cputime", C.c_double), ("walltime", C.c_double),
                      ("factorcputime", C.c_double), ("factorwalltime", C.c_double),
                      ("applycputime", C.c_double), ("applywalltime", C.c_double),
                      ("next", C.POINTER(ShellNode))]


class ShellList(C.Structure):
    _fields_ = [("ctxlist", C.POINTER(ShellNode)), ("size", C.c_int), ("factorcputime", C.c_double),
                ("factorwalltime", C.c_double), ("applycputime", C.c_double),
                ("applywalltime", C.c_double)]


_np = C.POINTER(ShellNode)
lib.b200_shell_list_new.restype = ShellList
lib.b200_shell_node_new.restype = ShellNode
lib.b200_shell_list_append.argtypes = [C.POINTER(ShellList), ShellNode]
lib.b200_shell_list_append.restype = None
lib.b200_shell_total_times.argtypes = [C.POINTER(ShellList)]
lib.b200_shell_total_times.restype = None
lib.b200_shell_list_destroy.argtypes = [C.POINTER(ShellList)]
lib.b200_shell_set_options.argtypes = [_np, C.POINTER(ShellOptions)]
lib.b200_shell_settings.argtypes = [_np, C.POINTER(Settings)]
lib.b200_shell_setup.argtypes = [_np, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
lib.b200_shell_apply.argtypes = [_np, C.c_void_p, C.c_void_p]
lib.b200_shell_apply_device.argtypes = [_np, C.c_void_p, C.c_void_p]
lib.b200_shell_relax.argtypes = [_np, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int,
                                 C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
lib.b200_shell_relax_device.argtypes = lib.b200_shell_relax.argtypes
lib.b200_shell_offers_relaxation.argtypes = [_np]
lib.b200_shell_cleanup.argtypes = [_np]
lib.b200_shell_info_count.argtypes = [_np]
lib.b200_shell_info_get.argtypes = [_np, C.c_int, C.c_void_p]


def make_options(pc_type: str, sweeps=(1, 1), scale=False, fact_init="init_original",
                 apply_init="init_jacobi", chunk=128, precinfo=False) -> ShellOptions:
    o = ShellOptions()
    o.pc_type = pc_type.encode()
    o.async_sweeps[0], o.async_sweeps[1] = int(sweeps[0]), int(sweeps[1])
    o.use_symmetric_scaling = int(scale)
    o.fact_init_type = fact_init.encode()
    o.apply_init_type = apply_init.encode()
    o.thread_chunk_size = chunk
    o.compute_preconditioner_info = int(precinfo)
    return o


class BlastedDataList:
    """Blasted_data_list: owns the nodes (and the device objects attached to them)."""

    def __init__(self):
        self.c = lib.b200_shell_list_new()

    def append_new(self) -> "BlastedData":
        """appendBlastedDataContext(list, newBlastedDataContext()); the new node is the head."""
        lib.b200_shell_list_append(C.byref(self.c), lib.b200_shell_node_new())
        # (a pointer-typed field read from a Structure aliases the field itself: take the address)
        return BlastedData(C.cast(C.addressof(self.c.ctxlist.contents), _np))

    def nodes(self):
        p = self.c.ctxlist
        while p:
            q = C.cast(C.addressof(p.contents), _np)
            yield BlastedData(q)
            p = q.contents.next

    def compute_total_times(self):
        lib.b200_shell_total_times(C.byref(self.c))
        return (self.c.factorwalltime, self.c.applywalltime, self.c.factorcputime, self.c.applycputime)

    def destroy(self) -> None:
        check(lib.b200_shell_list_destroy(C.byref(self.c)))


class BlastedData:
    """One Blasted_node, held by pointer into its list."""

    def __init__(self, ptr):
        self.p = ptr

    @property
    def node(self) -> ShellNode:
        return self.p.contents

    def set_options(self, opts: ShellOptions) -> None:
        check(lib.b200_shell_set_options(self.p, C.byref(opts)))

    def settings(self) -> Settings:
        s = Settings()
        check(lib.b200_shell_settings(self.p, C.byref(s)))
        return s

    def setup(self, m) -> None:
        """m: SRMatrix with column-major blocks (what PETSc's BAIJ stores); values are re-read."""
        di = m.diagind.ctypes.data_as(C.c_void_p) if m.diagind is not None else None
        check(lib.b200_shell_setup(self.p, m.bs, m.nbrows, m.browptr.ctypes.data_as(C.c_void_p),
                                   m.bcolind.ctypes.data_as(C.c_void_p),
                                   m.vals.ctypes.data_as(C.c_void_p), di))

    def apply(self, r, z=None):
        if type(r).__module__.startswith("torch"):
            import torch
            z = torch.empty_like(r) if z is None else z
            check(lib.b200_shell_apply_device(self.p, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr())))
            return z
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.zeros_like(r) if z is None else z
        check(lib.b200_shell_apply(self.p, r.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))
        return z

    def relax(self, rhs, x, its: int, guesszero: bool = False, rtol=0.0, abstol=0.0, dtol=0.0):
        outits, reason = C.c_int(), C.c_int()
        rhs = np.ascontiguousarray(rhs, dtype=np.float64)
        check(lib.b200_shell_relax(self.p, rhs.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                                   rtol, abstol, dtol, its, int(guesszero), C.byref(outits),
                                   C.byref(reason)))
        return outits.value, reason.value

    def offers_relaxation(self) -> bool:
        return bool(lib.b200_shell_offers_relaxation(self.p))

    def infos(self):
        out = []
        for i in range(lib.b200_shell_info_count(self.p)):
            rec = np.zeros(6)
            check(lib.b200_shell_info_get(self.p, i, rec.ctypes.data_as(C.c_void_p)))
            out.append(rec)
        return out

    def cleanup(self) -> None:
        check(lib.b200_shell_cleanup(self.p))
