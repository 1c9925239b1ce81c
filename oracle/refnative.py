"""CPU ORACLE (test infrastructure): ctypes bindings of oracle/_ref/libref_line_descriptor.so -- the reference's own
line_descriptor C++ (KeyLine fill, computeLBD, BinaryDescriptorMatcher) compiled unmodified by oracle/build_ref.py."""
import ctypes as C

import numpy as np

from . import build_ref

_lib = None


def lib():
    global _lib
    if _lib is None:
        p = build_ref.build()
        if p is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _lib = C.CDLL(p)
    return _lib


def available():
    try:
        lib()
        return True
    except Exception:
        return False


def keylines_lbd(gray, lines_px):
    """LSDDetectorC KeyLine fill on `lines_px` + BinaryDescriptor::compute -> (keylines f32 [S,10], desc72 f32 [S,72], desc32 u8 [S,32])."""
    gray = np.ascontiguousarray(gray, np.uint8)
    lines = np.ascontiguousarray(lines_px, np.float32).reshape(-1, 4)
    n = len(lines)
    H, W = gray.shape
    kl = np.zeros((n, 10), np.float32); d72 = np.zeros((n, 72), np.float32); d32 = np.zeros((n, 32), np.uint8)
    rc = lib().ref_keylines_lbd(gray.ctypes.data_as(C.c_void_p), H, W, lines.ctypes.data_as(C.c_void_p), n,
                                kl.ctypes.data_as(C.c_void_p), d72.ctypes.data_as(C.c_void_p), d32.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("ref_keylines_lbd failed: %d" % rc)
    return kl, d72, d32


def knn_match(q, m, k):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); m = np.ascontiguousarray(m, np.uint8).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32); dist = np.empty((len(q), k), np.int32)
    rc = lib().ref_knn_match(q.ctypes.data_as(C.c_void_p), len(q), m.ctypes.data_as(C.c_void_p), len(m), int(k),
                             idx.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
    if rc < 0:
        raise RuntimeError("ref_knn_match failed: %d" % rc)
    return idx, dist
