"""ctypes binding of libb200lrcn.so -- the C-ABI boundary declared in include/b200lrcn.h.

There is no CPU / other-GPU fallback: if the shared library is missing, or a call returns a
non-zero status, a B200LrcnError is raised."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_long, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lrcn.so")


class B200LrcnError(RuntimeError):
    pass


_P, _I, _L, _F = c_void_p, c_int, c_long, c_float

# name -> argument ctypes (all functions return int status unless listed in _RET)
SIGNATURES = {
    "b2_abi_version": [],
    "b2_last_error": [],
    "b2_launch_count": [],
    "b2_device_check": [],
    "b2_gemm_bf16_tn": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _I, _I, _P, _P, _P],
    "b2_conv2d_nhwc_bf16": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P],
}
_RET = {"b2_last_error": c_char_p, "b2_launch_count": c_long}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200LrcnError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = _RET.get(name, c_int)
        _lib = l
    return _lib


def call(name, *args):
    """Invoke an int-status entry point; raise with b2_last_error() on failure."""
    l = lib()
    rc = getattr(l, name)(*args)
    if rc != 0:
        msg = l.b2_last_error()
        raise B200LrcnError(f"{name} failed (status {rc}): {msg.decode() if msg else ''}")
    return rc


def launch_count() -> int:
    return int(lib().b2_launch_count())


def ptr(t):
    """data_ptr of a tensor or 0 for None."""
    return 0 if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
