"""ctypes binding of librvgp_b200.so (the C-ABI declared in include/rvgp_b200.h).

There is NO CPU fallback: if the library is missing, cannot be loaded, or no CUDA device is present
the product raises.  PyTorch is used by callers only to own device memory and streams; this module
passes raw pointers.
"""
import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librvgp_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "rvgp_b200.h")

RVGP_OK = 0
RVGP_ERR_BAD_ARG = -1
RVGP_ERR_RANK_DEFICIENT = -2
RVGP_ERR_CUDA = -3
RVGP_ERR_NCCL = -4
RVGP_ERR_NOT_SPD = -5
RVGP_ERR_CAPACITY = -6


class RvgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librvgp_b200 status %d: %s" % (code, msg))
        self.code = code


_lib = None
_lock = threading.Lock()


def declared_symbols():
    """Every function name declared in include/rvgp_b200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rvgp_[a-z0-9_]+)\s*\(", txt)))


def load_library():
    """dlopen the in-tree library.  Raises (never falls back) when it is absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "librvgp_b200.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "-- there is no CPU fallback." % LIB_PATH)
            lib = ctypes.CDLL(LIB_PATH)
            lib.rvgp_last_error.restype = ctypes.c_char_p
            lib.rvgp_launch_count.restype = ctypes.c_ulonglong
            for name in declared_symbols():
                if name.endswith("_workspace_bytes"):
                    getattr(lib, name).restype = ctypes.c_int64
            _lib = lib
    return _lib


def _conv(a):
    """Convert an argument: torch tensors -> device pointer; None -> NULL; python scalars by type."""
    if a is None:
        return ctypes.c_void_p(0)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, float):
        return ctypes.c_double(a)
    if isinstance(a, I64):
        return ctypes.c_int64(int(a))
    if isinstance(a, U64):
        return ctypes.c_uint64(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    if isinstance(a, (ctypes._SimpleCData, ctypes._Pointer, ctypes.Array)):
        return a
    raise TypeError("cannot pass %r to the C ABI" % type(a))


class I64(int):
    """Marks an int64_t argument."""


class U64(int):
    """Marks a uint64_t argument."""


class Handle:
    """One handle per (process, device).  Owns nothing but the C handle; follows torch's current stream."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("rvgp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = load_library()
        self.device = int(device)
        self._h = ctypes.c_void_p()
        rc = self.lib.rvgp_create(ctypes.c_int(self.device), ctypes.byref(self._h))
        if rc != 0:
            raise RvgpError(rc, "rvgp_create failed")
        self.sm_count = self.lib.rvgp_sm_count(self._h)

    def sync_stream(self):
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self.lib.rvgp_set_stream(self._h, ctypes.c_void_p(s))

    def call(self, name, *args):
        fn = getattr(self.lib, name)
        rc = fn(self._h, *[_conv(a) for a in args])
        if rc != 0:
            msg = self.lib.rvgp_last_error(self._h).decode("utf-8", "replace")
            if rc == RVGP_ERR_BAD_ARG:
                raise ValueError(msg)
            raise RvgpError(rc, msg)
        return rc

    def query(self, name, *args):
        return getattr(self.lib, name)(*[_conv(a) for a in args])

    def set_option(self, key, value):
        rc = self.lib.rvgp_set_option(self._h, ctypes.c_char_p(key.encode()), ctypes.c_int(int(value)))
        if rc != 0:
            raise ValueError(self.lib.rvgp_last_error(self._h).decode())

    @property
    def launches(self):
        return int(self.lib.rvgp_launch_count(self._h))

    def __del__(self):
        try:
            if self._h:
                self.lib.rvgp_destroy(self._h)
        except Exception:
            pass


_handles = {}


def get_handle(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    device = int(device)
    h = _handles.get(device)
    if h is None:
        h = Handle(device)
        _handles[device] = h
    h.sync_stream()
    return h
