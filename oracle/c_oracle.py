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
        _LIB.seir_oracle_joint_log_prob.restype = ctypes.c_int
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


def joint_log_prob(consts, car, initial_state, events, u, want_grad=False, num_threads=0):
    """joint_log_prob(unconstrained u [B,P], events [B,M,T,3]) of inference.py:537-557 -> [B] (and its gradient [B,P]);
    consts from seir_oracle.rate_constants, car from seir_oracle.car_constants."""
    lib = load()
    ev = np.ascontiguousarray(events, np.float64)
    uu = np.ascontiguousarray(u, np.float64)
    B, M, T, _ = ev.shape
    assert uu.shape == (B, 6 + (T - 1) + M)
    times = np.arange(T)
    W = np.ascontiguousarray(consts["W"][np.clip(times, 0, len(consts["W"]) - 1)])
    wk = np.ascontiguousarray(consts["weekday_c"][np.clip(times, 0, len(consts["weekday_c"]) - 1)])
    L = np.ascontiguousarray(car["scale_tril"], np.float64)
    out = np.empty(B, np.float64)
    grad = np.empty_like(uu) if want_grad else None
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else ctypes.c_void_p(0)
    keep = [np.ascontiguousarray(consts["Cstar"]), np.ascontiguousarray(consts["N"]), W, wk,
            np.ascontiguousarray(consts["log_area_c"]), np.ascontiguousarray(initial_state, np.float64)]
    rc = lib.seir_oracle_joint_log_prob(ctypes.c_int(B), ctypes.c_int(M), ctypes.c_int(T), *[p(a) for a in keep], p(L),
                                        ctypes.c_double(float(car["log_det_scale"])), p(ev), p(uu), ctypes.c_double(0.28),
                                        ctypes.c_double(1e-9), p(out), p(grad), ctypes.c_int(num_threads))
    if rc != 0:
        raise MemoryError("seir_oracle_joint_log_prob")
    return (out, grad) if want_grad else out


def max_threads():
    return int(load().seir_oracle_max_threads())
