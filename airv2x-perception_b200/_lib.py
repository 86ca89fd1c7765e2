"""ctypes loader for libairv2x_b200.so (the C-ABI in include/airv2x_b200.h).

The product path has no CPU fallback: if the shared library cannot be loaded this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libairv2x_b200.so")
_lib = None


class ConvShape(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("n", "h", "w", "cin", "cout", "ksize", "stride")]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build  # noqa: WPS433 (in-tree build; needs nvcc)

        _build.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libairv2x_b200.so is missing and could not be built; there is no CPU fallback")
    lib = ctypes.CDLL(LIB_PATH)
    lib.a2x_last_error.restype = ctypes.c_char_p
    lib.a2x_launch_count.restype = ctypes.c_ulonglong
    _lib = lib
    return lib


def _ptr(t):
    """torch tensor / int / None -> void* value."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    return ctypes.c_void_p(t.data_ptr())


PROFILE = None  # set to a list to record (name, args, start_event, end_event) per C-ABI call (bench roofline pass)


def call(name, *args):
    """Call a C-ABI entry point; raise RuntimeError with a2x_last_error() on a non-zero status."""
    lib = load()
    fn = getattr(lib, name)
    if PROFILE is not None:
        import torch

        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.append((name, args, e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.a2x_last_error().decode()))
    return rc


def stream_ptr():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
