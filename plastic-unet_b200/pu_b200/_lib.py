"""ctypes binding of the C-ABI kernel library (include/plastic_unet_b200.h).

The prototypes are parsed from the header itself so that the Python side can never drift from the
C-ABI.  There is NO fallback: if the shared library is missing or a call fails, we raise.
"""
from __future__ import annotations

import ctypes
import os
import re
from pathlib import Path

_HERE = Path(__file__).resolve().parent
REPO_ROOT = _HERE.parent.parent
HEADER = REPO_ROOT / "include" / "plastic_unet_b200.h"
LIB_PATH = Path(os.environ.get("PU_B200_LIB", _HERE / "lib" / "libpu_b200.so"))

_CTYPE = {
    "const float*": ctypes.c_void_p,
    "float*": ctypes.c_void_p,
    "const double*": ctypes.c_void_p,
    "const long long*": ctypes.c_void_p,
    "double*": ctypes.c_void_p,
    "void*": ctypes.c_void_p,
    "int*": ctypes.c_void_p,
    "const int*": ctypes.c_void_p,
    "unsigned char*": ctypes.c_void_p,
    "const unsigned char*": ctypes.c_void_p,
    "double": ctypes.c_double,
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
}
_RET = {"int": ctypes.c_int, "long long": ctypes.c_longlong, "void": None, "const char*": ctypes.c_char_p}


def parse_header(path: Path = HEADER):
    """-> {name: (restype_str, [(ctype_str, argname), ...])} for every function the header declares."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"#[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"(int|long long|void|const char\*)\s+(pu_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.+?)\s*(\w+)$", a)
                ty = mm.group(1).replace(" *", "*").strip()
                parsed.append((ty, mm.group(2)))
        protos[name] = (ret, parsed)
    return protos


class PuError(RuntimeError):
    pass


_lib = None
_protos = None


def load():
    """Load the kernel library (once).  Raises if it has not been built — never falls back."""
    global _lib, _protos
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise PuError(
            f"plastic-unet_b200 kernel library not found at {LIB_PATH}; build it with "
            f"`make -C {REPO_ROOT / 'plastic-unet_b200' / 'csrc'}` or `python -c 'import __graft_entry__ as g; g.build()'`"
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    protos = parse_header()
    for name, (ret, args) in protos.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = _RET[ret]
        fn.argtypes = [_CTYPE[t] for t, _ in args]
    _lib, _protos = lib, protos
    return lib


def call(name: str, *args):
    """Call an int-returning pu_* entry point; raise PuError(pu_last_error()) on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise PuError(f"{name} failed (status {rc}): {lib.pu_last_error().decode(errors='replace')}")


def launch_count() -> int:
    return int(load().pu_launch_count())


def reset_launch_count() -> None:
    load().pu_reset_launch_count()


def tc_available() -> bool:
    return bool(load().pu_tc_available())


def version() -> int:
    return int(load().pu_version())
