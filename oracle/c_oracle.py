"""ctypes loader of oracle/libseir_oracle.so (the C restatement) -- test/baseline infrastructure only."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return os.path.join(HERE, "libseir_oracle.so")


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "libseir_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.seir_oracle_log_prob.restype = ctypes.c_int
        _LIB.seir_oracle_max_threads.restype = ctypes.c_int
    return _LIB


def log_prob(consts, initial_state, events, theta, num_threads=0):
    """events [B,M,T,3], theta [B,P] (constrained) -> seir log-prob [B]; consts from seir_oracle.rate_constants."""
    lib = load()
    ev = np.ascontiguousarray(events, np.float64)
    th = np.ascontiguousarray(theta, np.float64)
    B, M, T, _ = ev.shape
    times = np.arange(T)
    W = np.ascontiguousarray(consts["W"][np.clip(times, 0, len(consts["W"]) - 1)])
    wk = np.ascontiguousarray(consts["weekday_c"][np.clip(times, 0, len(consts["weekday_c"]) - 1)])
    arrs = [np.ascontiguousarray(consts["Cstar"]), np.ascontiguousarray(consts["N"]), W, wk,
            np.ascontiguousarray(consts["log_area_c"]), np.ascontiguousarray(initial_state, np.float64), ev, th]
    out = np.empty(B, np.float64)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.seir_oracle_log_prob(ctypes.c_int(B), ctypes.c_int(M), ctypes.c_int(T), *[p(a) for a in arrs],
                                  ctypes.c_double(0.28), ctypes.c_double(1e-9), p(out), ctypes.c_int(num_threads))
    if rc != 0:
        raise MemoryError("seir_oracle_log_prob")
    return out


def max_threads():
    return int(load().seir_oracle_max_threads())
