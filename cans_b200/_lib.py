"""ctypes binding of the C ABI in include/cans_b200.h.

The shared library is built in-tree by `__graft_entry__.build()`
(`cans_b200/lib/libcans_b200.so`).  There is NO fallback: if the library is
missing or a symbol does not resolve, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcans_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1


class Options(C.Structure):
    _fields_ = [("thomas_variant", C.c_int), ("cache_slots", C.c_int), ("fft_x_lines", C.c_int),
                ("fft_y_lines", C.c_int), ("exchange", C.c_int), ("lambda_order", C.c_int), ("pivot_dedup", C.c_int),
                ("tall_tiles", C.c_int), ("reserved", C.c_int * 8)]

    def __init__(self, **kw):
        super().__init__()
        for name, _ in self._fields_[:-1]:
            setattr(self, name, kw.pop(name, -1))
        for i in range(8):
            self.reserved[i] = -1
        if kw:
            raise TypeError(f"unknown option(s): {sorted(kw)}")


# every symbol include/cans_b200.h declares: name -> (restype, argtypes)
_I3 = C.POINTER(C.c_int)
_D3 = C.POINTER(C.c_double)
_VP = C.c_void_p
SYMBOLS = {
    "cansb200_init": (C.c_int, [C.POINTER(_VP), _I3, _I3, C.c_int, C.c_int, C.c_int, _VP, C.c_int]),
    "cansb200_finalize": (C.c_int, [_VP]),
    "cansb200_dist_blob_size": (C.c_int, []),
    "cansb200_dist_export": (C.c_int, [_VP, _VP]),
    "cansb200_dist_connect": (C.c_int, [_VP, _VP]),
    "cansb200_dist_status": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "cansb200_dist_connect_local": (C.c_int, [C.POINTER(_VP), C.c_int]),
    "cansb200_get_extents": (C.c_int, [_VP, _I3, _I3, _I3, _I3]),
    "cansb200_plan_create": (C.c_int, [_VP, C.POINTER(_VP), C.c_char_p, C.c_char_p, C.POINTER(Options), _D3]),
    "cansb200_plan_destroy": (C.c_int, [_VP]),
    "cansb200_solve": (C.c_int, [_VP, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, _VP, C.c_int, _VP]),
    "cansb200_solve_z": (C.c_int, [_VP, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, C.c_int, _VP]),
    "cansb200_solve_z_bc": (C.c_int, [_VP, C.c_char_p, C.c_char, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, C.c_int, _VP]),
    "cansb200_fftini": (C.c_int, [_VP, C.c_char_p, C.c_char_p, _D3, C.POINTER(C.c_int)]),
    "cansb200_fftend": (C.c_int, [_VP, C.c_int]),
    "cansb200_solver": (C.c_int, [_VP, C.c_int, C.c_char_p, C.c_char_p, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, _VP, C.c_int,
                                  C.c_int, _VP]),
    "cansb200_solve_fillps": (C.c_int, [_VP, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, _VP, _D3, _VP, C.c_double, _VP, _VP, _VP,
                                        _I3, _I3, _D3, _VP]),
    "cansb200_solver_fillps": (C.c_int, [_VP, C.c_int, C.c_char_p, C.c_char_p, _VP, _I3, C.c_int, C.c_double, _VP, _VP, _VP, _VP,
                                         C.c_int, _D3, _VP, C.c_double, _VP, _VP, _VP, _I3, _I3, _D3, _VP]),
    "cansb200_plan_id": (C.c_int, [_VP]),
    "cansb200_plan_from_id": (_VP, [C.c_int]),
    "cansb200_updt_rhs_b": (C.c_int, [_VP, C.c_char_p, C.c_char_p, _I3, _I3, _I3, _D3, C.c_double, _VP, _VP]),
    "cansb200_r2r": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _I3, _VP]),
    "cansb200_gaussel": (C.c_int, [_VP, _VP, _I3, C.c_int, C.c_int, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "cansb200_gaussel_dtdma": (C.c_int, [_VP, _VP, _I3, C.c_int, C.c_int, _I3, C.c_int, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "cansb200_fillps": (C.c_int, [_VP, _I3, _D3, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "cansb200_correc": (C.c_int, [_VP, _I3, _D3, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "cansb200_chkdiv": (C.c_int, [_VP, _I3, _D3, _VP, _VP, _VP, _VP, _D3, _D3, _VP]),
    "cansb200_fill_hash": (C.c_int, [_VP, _VP, _I3, _I3, C.c_int, C.c_ulonglong, _VP]),
    "cansb200_last_error": (C.c_char_p, []),
    "cansb200_version": (C.c_int, []),
    "cansb200_plan_stats": (C.c_int, [_VP, C.POINTER(C.c_ulonglong)]),
    "cansb200_set_profiling": (C.c_int, [_VP, C.c_int]),
    "cansb200_get_profile": (C.c_int, [_VP, _D3, C.POINTER(C.c_ulonglong)]),
    "cansb200_ctx_set": (C.c_int, [_VP, C.c_int, C.c_int]),
    "cansb200_get_work": (C.c_int, [_VP, C.c_int, C.POINTER(_VP), C.POINTER(C.c_size_t)]),
}


def load(path: str = LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(f"cans_b200: {path} not found -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: intended
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class CansError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib.cansb200_last_error()
        raise CansError(f"{what}: status {rc}: {msg.decode() if msg else ''}")


def i3(v):
    return (C.c_int * len(v))(*[int(x) for x in v])


def d3(v):
    return (C.c_double * len(v))(*[float(x) for x in v])
