"""ctypes binding of libspaa_b200.so, generated from include/spaa_b200.h.

Every prototype in the header is parsed and bound, so the header is the single source of truth for the C ABI.
There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "spaa_b200.h")
LIB_PATH = os.path.join(HERE, "libspaa_b200.so")

_SCALARS = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float,
            "spaa_stream_t": ctypes.c_void_p}


class ConvDesc(ctypes.Structure):
    """Mirror of `spaa_conv_desc` (include/spaa_b200.h)."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("in_dtype", "out_dtype", "B", "Cin", "Hin", "Win", "Cout", "Hout", "Wout",
                                               "KH", "KW", "stride", "up", "pad_h", "pad_w", "flip")]
                + [(n, ctypes.c_int64) for n in ("in_bs", "in_ps", "in_cs", "w_ts", "w_cis", "w_cos", "out_bs", "out_ps",
                                                 "out_cs", "add_bs", "add_ps", "add_cs", "mask_bs", "mask_ps", "mask_cs")]
                + [(n, ctypes.c_int32) for n in ("epi_flags", "mask_mode", "split")])


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """Returns {function name: (return type, [argument C types])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"typedef struct spaa_conv_desc\s*\{.*?\}\s*spaa_conv_desc\s*;", " ", src, flags=re.S)
    src = re.sub(r"enum\s*\{.*?\}\s*;", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\s*\*|int64_t|int)\s+(spaa_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        argl = [] if args.strip() in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (re.sub(r"\s+", " ", ret.strip()), [re.sub(r"\s+", " ", a) for a in argl])
    return protos


def _ctype(decl: str):
    if "*" in decl:
        return ctypes.c_void_p
    t = decl.replace("const ", "").split()[0]
    return _SCALARS[t]


class SpaaError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python spaa_b200/build.py` "
                              "(or __graft_entry__.build()); spaa_b200 has no CPU / PyTorch fallback")
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)          # AttributeError if the library does not export a declared symbol
            fn.argtypes = [_ctype(a) for a in args]
            fn.restype = ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if ret == "int64_t" else ctypes.c_int)
            if ret == "int" and name != "spaa_abi_version" and not name.endswith("_supported"):
                setattr(self, name, self._checked(name, fn))     # status-returning entry point: raise on non-zero
            else:
                setattr(self, name, fn)

    def _checked(self, name, fn):
        last_error = self.cdll.spaa_last_error

        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise SpaaError(f"{name} failed ({rc}): {last_error().decode()}")
            return rc
        return call


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
