"""ctypes access to oracle/liboracle.so (the C restatement).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(_PATH)
        l.oracle_xxh64.restype = C.c_uint64
        l.oracle_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        l.oracle_chunk_keys.restype = C.c_int
        l.oracle_chunk_keys.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_void_p]
        l.oracle_num_threads.restype = C.c_int
        l.oracle_set_threads.argtypes = [C.c_int]
        for name in ("oracle_gather_raw", "oracle_scatter_raw"):
            f = getattr(l, name)
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_uint32, C.c_void_p, C.c_int64,
                          C.c_int, C.c_void_p, C.c_uint64]
        for name in ("oracle_gather_fp8", "oracle_scatter_fp8"):
            f = getattr(l, name)
            f.restype = None
            f.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                          C.c_int, C.c_void_p, C.c_uint64, C.c_uint64]
        _lib = l
    return _lib


def planes_of(layers):
    """layers: list of C-contiguous uint16 arrays (2, NB, bs, H, D) -> (void*[2L], block_stride)."""
    ptrs = []
    for a in layers:
        assert a.flags["C_CONTIGUOUS"]
        ptrs += [a[0].ctypes.data, a[1].ctypes.data]
    arr = (C.c_void_p * len(ptrs))(*ptrs)
    two, nb, bs, h, d = layers[0].shape
    return arr, bs * h * d * layers[0].itemsize


def gather(layers, slot_mapping, chunk_tokens, fmt="raw"):
    two, nb, bs, h, d = layers[0].shape
    planes, stride = planes_of(layers)
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    n = len(sm)
    nch = (n + chunk_tokens - 1) // chunk_tokens
    L2 = 2 * len(layers)
    if fmt == "raw":
        cb = L2 * chunk_tokens * h * d * 2
        out = np.zeros(nch * cb, dtype=np.uint8)
        lib().oracle_gather_raw(planes, L2, stride, bs, h * d * 2, sm.ctypes.data, n, chunk_tokens,
                                out.ctypes.data, cb)
        return out, cb, 0
    so = L2 * chunk_tokens * h * d
    cb = (so + L2 * h * 4 + 255) // 256 * 256
    out = np.zeros(nch * cb, dtype=np.uint8)
    lib().oracle_gather_fp8(planes, L2, stride, bs, h, d, sm.ctypes.data, n, chunk_tokens,
                            out.ctypes.data, cb, so)
    return out, cb, so


def scatter(layers, slot_mapping, chunk_tokens, chunks, cb, so, fmt="raw"):
    two, nb, bs, h, d = layers[0].shape
    planes, stride = planes_of(layers)
    sm = np.ascontiguousarray(slot_mapping, dtype=np.int64)
    L2 = 2 * len(layers)
    if fmt == "raw":
        lib().oracle_scatter_raw(planes, L2, stride, bs, h * d * 2, sm.ctypes.data, len(sm), chunk_tokens,
                                 chunks.ctypes.data, cb)
    else:
        lib().oracle_scatter_fp8(planes, L2, stride, bs, h, d, sm.ctypes.data, len(sm), chunk_tokens,
                                 chunks.ctypes.data, cb, so)
