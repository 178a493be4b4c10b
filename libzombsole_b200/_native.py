"""Loader of the CUDA shared library (csrc/libzs_b200.so) over ctypes.

There is NO fallback: if the library has not been built, or a call fails, this raises.
Build it with ``python -m libzombsole_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

from . import abi

#: ZS_B200_LIB points at another build of the same library (development: comparing two builds on one box)
LIB_PATH = os.environ.get("ZS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libzs_b200.so")
_lib = None


class ZsError(RuntimeError):
    pass


def lib():
    """The loaded library with every prototype of include/zs_b200.h bound."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ZsError("CUDA extension %s is missing: build it with `python -m libzombsole_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in abi.PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if L.zs_abi_version() != abi.ZS_ABI_VERSION:
            raise ZsError("ABI version mismatch: library %d, python %d" % (L.zs_abi_version(), abi.ZS_ABI_VERSION))
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise ZsError(lib().zs_last_error().decode("utf-8", "replace"))
