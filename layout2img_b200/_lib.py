"""ctypes binding of libl2i.so (the C ABI in include/l2i.h).

The library is the product: importing an op without it raises -- there is no eager/CPU
fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libl2i.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "l2i.h")

_lib = None


class L2IError(RuntimeError):
    pass


def declared_symbols():
    """Every function name include/l2i.h declares (used by the export test)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(l2i_[a-z0-9_]+)\s*\(", txt)))


_CTYPES = {"int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
           "long long": ctypes.c_longlong, "int64_t": ctypes.c_int64}


def prototypes():
    """{name: [ctypes argument types]} parsed from include/l2i.h (pointers -> c_void_p)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(l2i_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        name, args = m.group(1), m.group(2).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(ctypes.c_void_p)
                else:
                    base = re.sub(r"\bconst\b", "", a).strip().rsplit(" ", 1)[0].strip()
                    types.append(_CTYPES[base])
        out[name] = types
    return out


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise L2IError(
                f"{LIB_PATH} is missing: build it with `python -m layout2img_b200.build` "
                "(nvcc, sm_100a). There is no fallback path.")
        _lib = ctypes.CDLL(LIB_PATH)
        for name, types in prototypes().items():
            fn = getattr(_lib, name)
            fn.argtypes = types
            fn.restype = ctypes.c_char_p if name == "l2i_last_error" else ctypes.c_int
    return _lib


def _conv(a):
    import torch
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    if isinstance(a, bool):
        return int(a)
    return a


def call(name: str, *args):
    """Call an int-returning l2i_* entry point; tensors become raw device pointers, the
    current torch CUDA stream is appended as the trailing `void* stream` argument."""
    import torch
    fn = getattr(lib(), name)
    rc = fn(*[_conv(a) for a in args], torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        msg = lib().l2i_last_error().decode(errors="replace")
        if rc in (-1, -2):
            raise ValueError(f"{name}: {msg}")
        raise L2IError(f"{name} failed ({rc}): {msg}")
