"""ctypes binding of librt_wire.so (include/rt_wire.h): the master-side decoder of comms.TraceResults -- plain C, no CUDA.
Mirrors what the reference master does with a reply (master/main.go:166-176: GetResults() + surface.Set per pixel)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librt_wire.so")
WIRE_SYMBOLS = ("rt_wire_decode_trace_results", "rt_wire_blit_trace_results")
ERR_MALFORMED, ERR_CAPACITY, ERR_ARGUMENT = -1, -2, -3
_LIB = None


class WireError(ValueError):
    def __init__(self, code):
        super().__init__({-1: "malformed TraceResults message", -2: "more results than expected", -3: "bad argument"}.get(code, "error %d" % code))
        self.code = code


def load():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        L.rt_wire_decode_trace_results.restype = C.c_int64
        L.rt_wire_decode_trace_results.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.rt_wire_blit_trace_results.restype = C.c_int
        L.rt_wire_blit_trace_results.argtypes = [C.c_void_p, C.c_uint64] + [C.c_uint32] * 4 + [C.c_void_p] + [C.c_uint32] * 7
        _LIB = L
    return _LIB


def _buf(wire):
    a = np.frombuffer(bytes(wire), dtype=np.uint8) if not isinstance(wire, np.ndarray) else np.ascontiguousarray(wire, dtype=np.uint8)
    return a, (a.ctypes.data if a.size else None)


def decode_trace_results(wire, max_results=None):
    """-> (n, 3) uint8 array of the results in message order."""
    a, p = _buf(wire)
    cap = a.size // 2 + 1 if max_results is None else int(max_results)       # an element is at least 2 bytes
    out = np.empty((cap, 3), dtype=np.uint8)
    n = load().rt_wire_decode_trace_results(p, a.size, out.ctypes.data, cap)
    if n < 0:
        raise WireError(int(n))
    return out[:n]


def blit_trace_results(wire, x, y, w, h, surface, shifts=(16, 8, 0), amask=0):
    """surface: (H, W) uint32 array (a 32-bit window surface); writes region (x, y, w, h) of it from one order's reply."""
    a, p = _buf(wire)
    assert surface.dtype == np.uint32 and surface.ndim == 2 and surface.strides[1] == 4
    rc = load().rt_wire_blit_trace_results(p, a.size, x, y, w, h, surface.ctypes.data, surface.shape[1], surface.shape[0],
                                           surface.strides[0], shifts[0], shifts[1], shifts[2], amask)
    if rc != 0:
        raise WireError(rc)
    return surface
