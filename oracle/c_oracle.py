"""ctypes wrapper of oracle/libcheb_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcheb_oracle.so")


class layer64_t(C.Structure):
    _fields_ = [("K", C.c_int), ("f_in", C.c_int), ("f_out", C.c_int), ("act", C.c_int), ("slope", C.c_double),
                ("W", C.c_void_p), ("b", C.c_void_p)]


def build(force=False):
    src = os.path.join(HERE, "cheb_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libcheb_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.cheb_stack_forward_f64.restype = C.c_int
        _lib.cheb_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().cheb_oracle_max_threads())


def stack_forward(graph_off, rowptr, colidx, vals, weights, acts, slope, X, n_threads=0):
    """fp64, one graph at a time, graphs spread over n_threads OpenMP threads (0 = all)."""
    graph_off = np.ascontiguousarray(graph_off, dtype=np.int32)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    colidx = np.ascontiguousarray(colidx, dtype=np.int32)
    X = np.ascontiguousarray(X, dtype=np.float64)
    vals = None if vals is None else np.ascontiguousarray(vals, dtype=np.float64)
    keep, arr = [], (layer64_t * len(weights))()
    for i, ((W, b), act) in enumerate(zip(weights, acts)):
        W = np.ascontiguousarray(W, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        keep += [W, b]
        arr[i].K, arr[i].f_in, arr[i].f_out, arr[i].act, arr[i].slope = W.shape[0], W.shape[1], W.shape[2], act, slope
        arr[i].W, arr[i].b = W.ctypes.data, b.ctypes.data
    Y = np.empty((X.shape[0], weights[-1][0].shape[2]), dtype=np.float64)
    rc = lib().cheb_stack_forward_f64(C.c_int(graph_off.size - 1), C.c_void_p(graph_off.ctypes.data),
                                      C.c_void_p(rowptr.ctypes.data), C.c_void_p(colidx.ctypes.data),
                                      C.c_void_p(vals.ctypes.data if vals is not None else None), arr,
                                      C.c_int(len(weights)), C.c_void_p(X.ctypes.data), C.c_void_p(Y.ctypes.data),
                                      C.c_int(int(n_threads)))
    if rc != 0:
        raise MemoryError("cheb_stack_forward_f64 failed")
    return Y
