"""ctypes binding of libdsa.so (include/dsa.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in the product path."""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdsa.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "dsa.h")

DSA_OK, DSA_ERR_ARGUMENT, DSA_ERR_BOUNDS, DSA_ERR_ERROR = 0, 1, 2, 3
DSA_ERR_CUDA, DSA_ERR_OOM, DSA_ERR_INTERNAL = 10, 11, 12
COMBINE = {"+": 0, "add": 0, "*": 1, "mul": 1, "last": 2, "first": 3, "min": 4, "max": 5}
COLMAJOR, ROWMAJOR = 0, 1


class DsaError(Exception):
    """Base of the exceptions raised by the library (code = DSA_ERR_*)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class ArgumentError(DsaError, ValueError):      # Julia ArgumentError
    pass


class BoundsError(DsaError, IndexError):        # Julia BoundsError
    pass


class ErrorException(DsaError, RuntimeError):   # Julia ErrorException (error("..."))
    pass


class CudaError(DsaError, RuntimeError):
    pass


_EXC = {DSA_ERR_ARGUMENT: ArgumentError, DSA_ERR_BOUNDS: BoundsError, DSA_ERR_ERROR: ErrorException,
        DSA_ERR_CUDA: CudaError, DSA_ERR_OOM: CudaError, DSA_ERR_INTERNAL: DsaError}


def build(force=False, verbose=False):
    """Compile libdsa.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [_HEADER]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", csrc, "../libdsa.so"], stdout=None if verbose else subprocess.DEVNULL)
    return _SO


def declared_symbols():
    """Every function name include/dsa.h declares."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsa_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(
                f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libdsa has no CPU fallback; the CUDA library is the product)")
        L = C.CDLL(_SO)
        L.dsa_last_error.restype = C.c_char_p
        for name in ("dsa_spread_dest", "dsa_spread_rank", "dsa_colmap_plan", "dsa_launch_count", "dsa_prof_dump"):
            getattr(L, name).restype = C.c_int64
        _lib = L
    return _lib


def check(code):
    if code != 0:
        msg = lib().dsa_last_error().decode(errors="replace")
        raise _EXC.get(code, DsaError)(code, msg)


def set_tile_mode(mode):
    """How a batched setindex! is applied (include/dsa.h, "tuning"): 0 = random-access pipeline only, 1 = dense batches are
    tile-streamed (default), 2 = tile-streamed whenever the structure allows it.  Same layout either way.  Returns the previous mode."""
    return int(lib().dsa_set_tile_mode(C.c_int(int(mode))))


def device_count():
    n = C.c_int()
    lib().dsa_device_count(C.byref(n))
    return n.value


def require_gpu():
    if device_count() == 0:
        raise CudaError(DSA_ERR_CUDA, "no CUDA device available: libdsa has no CPU fallback")
