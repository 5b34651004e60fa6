"""ctypes binding of libb200lrcn.so -- the C-ABI boundary declared in include/b200lrcn.h.

The header is the single source of truth: its prototypes are parsed here to build the ctypes
signatures, so the Python side cannot drift from the ABI.  There is no CPU / other-GPU fallback:
if the shared library is missing, or a call returns a non-zero status, B200LrcnError is raised."""
import ctypes
import os
import re
from ctypes import c_char_p, c_float, c_int, c_long, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2_LIB_OVERRIDE") or os.path.join(_HERE, "libb200lrcn.so")   # override: A/B kernel experiments
HEADER_PATH = os.path.join(_HERE, "..", "include", "b200lrcn.h")


class B200LrcnError(RuntimeError):
    pass


_SCALARS = {"int": c_int, "long": c_long, "float": c_float, "unsigned": c_ulonglong}


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every `b2_*` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for ret, name, args in re.findall(r"\b(int|long|const char\*)\s+(b2_\w+)\s*\(([^)]*)\)\s*;", src):
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(c_void_p)
                else:
                    argtypes.append(_SCALARS[a.split()[0]])
        restype = {"int": c_int, "long": c_long, "const char*": c_char_p}[ret]
        out[name] = (restype, argtypes)
    return out


_lib = None
_sigs = None


def lib():
    global _lib, _sigs
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200LrcnError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no fallback path)")
        l = ctypes.CDLL(LIB_PATH)
        _sigs = parse_header()
        for name, (restype, argtypes) in _sigs.items():
            fn = getattr(l, name)      # AttributeError if the .so is stale w.r.t. the header
            fn.argtypes = argtypes
            fn.restype = restype
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


_device_ok = False


def require_device():
    """Fail loudly unless the current CUDA device is a B200 (sm_100)."""
    global _device_ok
    if not _device_ok:
        import torch
        if not torch.cuda.is_available():
            raise B200LrcnError("b200-lrcn needs a CUDA device (sm_100a); no CPU fallback exists")
        call("b2_device_check")
        _device_ok = True


def ptr(t):
    """data_ptr of a tensor, a raw device address (int) as is, or 0 (NULL) for None."""
    if t is None:
        return 0
    return t if isinstance(t, int) else t.data_ptr()


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (the raw C getters: this is called once per
    kernel launch, torch.cuda.current_stream() costs ~15 us of Python per call)."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        _raw_stream = (torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice)
    return _raw_stream[0](_raw_stream[1]())
