"""ctypes binding of libbvae_b200.so (the C ABI declared in include/bvae_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing, or a
symbol the header declares is absent, importing a kernel raises immediately."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "libbvae_b200.so")
HEADER_PATH = os.path.join(REPO_ROOT, "include", "bvae_b200.h")

F32, BF16 = 0, 1
LOSS = {"mse": 0, "l1": 1, "bce": 2}

vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

def _parse_header():
    """Derive every ctypes signature from the header itself, so the binding cannot drift."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    sigs = {}
    for m in re.finditer(r"(const char\*|size_t|int|unsigned long long)\s+(bvae_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        res = {"const char*": C.c_char_p, "size_t": sz, "int": i32, "unsigned long long": C.c_ulonglong}[ret]
        args = []
        for prm in [q.strip() for q in params.split(",")]:
            if prm in ("void", ""):
                continue
            if "*" in prm or "bvae_stream_t" in prm:
                args.append(vp)
            elif prm.startswith("size_t"):
                args.append(sz)
            elif prm.startswith("float"):
                args.append(f32)
            elif prm.startswith("int"):
                args.append(i32)
            else:
                raise RuntimeError(f"cannot map parameter '{prm}' of {name}")
        sigs[name] = (res, args)
    return sigs


_SIGS = _parse_header()

_lib = None


def header_symbols():
    """Every function name include/bvae_b200.h declares."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bvae_[a-z0-9_]+)\s*\(", txt)))


def load():
    """dlopen the library and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "This package has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name in header_symbols():
        if name not in _SIGS:
            raise RuntimeError(f"header declares {name} but bvae_rt.lib has no signature for it")
    for name, (res, args) in _SIGS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().bvae_last_error().decode("utf-8", "replace")


class BvaeError(RuntimeError):
    pass


def call(name, *args):
    """Invoke an int-returning entry point; non-zero return -> BvaeError with the library's message."""
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise BvaeError(f"{name} failed ({rc}): {last_error()}")


def query(name, *args) -> int:
    return int(getattr(load(), name)(*args))
