"""The C-ABI library loads without a GPU and exports every symbol include/cans_b200.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "cans_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cansb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from cans_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/cans_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_argument_errors_do_not_abort():
    """Error convention: status codes, never abort (the reference `error stop`s)."""
    from cans_b200 import _lib
    lib = _lib.lib
    assert lib.cansb200_version() >= 100
    assert lib.cansb200_init(None, None, None, 1, 0, 1, None, 0) == -1  # CANSB200_EINVAL
    assert b"null" in lib.cansb200_last_error()
    h = C.c_void_p()
    assert lib.cansb200_init(C.byref(h), _lib.i3([0, 4, 4]), _lib.i3([1, 1]), 1, 0, 1, None, 0) == -1
    assert lib.cansb200_init(C.byref(h), _lib.i3([4, 4, 4]), _lib.i3([1, 1]), 4, 0, 1, None, 0) == -1  # no such pencil axis
    assert lib.cansb200_init(C.byref(h), _lib.i3([4, 4, 4]), _lib.i3([1, 2]), 3, 0, 2, None, 0) == -4  # z pencils of a decomposed grid
    assert lib.cansb200_solve(None, None, None, 1, 1.0, None, None, None, None, 1, None) == -1
    assert lib.cansb200_plan_destroy(None) == 0
    assert lib.cansb200_finalize(None) == 0


def test_product_package_does_not_import_oracle():
    """The shipped path must not route through the CPU oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cans_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle.cans_oracle.hash_field", "") \
                    or f in ("solver.py",) and "import oracle" not in src and "from oracle" not in src, f
