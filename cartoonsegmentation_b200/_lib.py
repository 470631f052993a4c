"""ctypes binding of libcsb200.so (the C-ABI CUDA library, include/csb200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("CSB200_LIB") or os.path.join(_HERE, "libcsb200.so")        # CSB200_LIB: an A/B build of the same library (build.py)
_lib = None


class CsbError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise CsbError(f"{_SO} is missing: build it with `python -m cartoonsegmentation_b200.build` "
                           "(there is no CPU fallback for this path)")
        _lib = C.CDLL(_SO)
        _lib.csb_last_error.restype = C.c_char_p
        _lib.csb_launch_count.restype = C.c_uint64
    return _lib


def check(status, what=""):
    if status != 0:
        raise CsbError(f"{what} failed with status {status}: {lib().csb_last_error().decode()}")


def launch_count():
    return int(lib().csb_launch_count())


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CsbError("expected a CUDA tensor: this path has no CPU implementation")
    if not t.is_contiguous():
        raise CsbError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def f3(v):
    return (C.c_float * 3)(float(v[0]), float(v[1]), float(v[2]))
