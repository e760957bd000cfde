"""ctypes binding of the C ABI in include/gnnepcsaft_b200.h.

There is no CPU or library fallback: if the shared library is missing and cannot be built,
or a kernel returns non-zero, a RuntimeError is raised.
"""

from __future__ import annotations

import ctypes
import os
import re
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_PKG, "..", "csrc"))
# GNN_B200_LIB: load another build of the same ABI (A/B timing of kernel variants, the -DGNN_B200_PROFILE build)
LIB_PATH = os.environ.get("GNN_B200_LIB") or os.path.join(CSRC, "libgnnepcsaft_b200.so")

HEADER_PATH = os.path.normpath(os.path.join(_PKG, "..", "..", "include", "gnnepcsaft_b200.h"))

_CT = {"p": ctypes.c_void_p, "l": ctypes.c_int64, "i": ctypes.c_int, "f": ctypes.c_float}
_DECL = re.compile(r"^(int64_t|int|const char\*)\s+(gnn_\w+)\s*\(([^)]*)\)\s*;", re.M | re.S)


def _parse_header():
    """The header is the single source of truth: (name -> (return kind, arg kinds))."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for ret, name, args in _DECL.findall(text):
        kinds = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds.append("p")
                elif a.startswith("int64_t"):
                    kinds.append("l")
                elif a.startswith(("int32_t", "int ")):
                    kinds.append("i")
                elif a.startswith("float"):
                    kinds.append("f")
                else:
                    raise RuntimeError(f"unparsed argument '{a}' of {name} in {HEADER_PATH}")
        sigs[name] = (ret, "".join(kinds))
    return sigs


_lib = None
_lock = threading.Lock()


def exported_symbols():
    """Every symbol include/gnnepcsaft_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(_parse_header().keys())


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise RuntimeError(f"gnnepcsaft_b200: {LIB_PATH} is missing (run __graft_entry__.build())")
            import importlib.util
            spec = importlib.util.spec_from_file_location("_gnn_b200_build", os.path.join(CSRC, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (ret, sig) in _parse_header().items():
            fn = getattr(lib, name)
            fn.restype = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "const char*": ctypes.c_char_p}[ret]
            fn.argtypes = [_CT[ch] for ch in sig]
        _lib = lib
    return _lib


def _arg(a):
    if isinstance(a, torch.Tensor):
        return a.data_ptr()
    if a is None:
        return None
    return a


def call(name: str, *args):
    """Invoke an int-returning entry point; tensors are passed as raw pointers; raises on error."""
    lib = load()
    rc = getattr(lib, name)(*[_arg(a) for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.gnn_error_string(rc).decode()} (code {rc})")


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream
