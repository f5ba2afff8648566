"""ctypes binding of ``include/seir_b200.h`` (the C-ABI drop-in boundary).

There is no CPU fallback: if ``libseir_b200.so`` is missing or fails to load, importing the compute
path raises.  Build it in-tree with ``python -m covid19uk_b200.build``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

from . import build as _build

THETA_CONSTRAINED = 0
THETA_UNCONSTRAINED = 1
PART_SEIR = 1
PART_PRIORS = 2
PART_ILDJ = 4
PART_JOINT = 7
ABI_VERSION = 10


class SeirSpec(ctypes.Structure):
    _fields_ = [
        ("num_meta", c_int32),
        ("num_steps", c_int32),
        ("initial_step", c_int32),
        ("n_commute_volume", c_int32),
        ("n_weekday", c_int32),
        ("car_nnz", c_int32),
        ("time_delta", c_double),
        ("nu", c_double),
        ("rate_eps", c_double),
        ("car_log_det_scale", c_double),
        ("cstar", POINTER(c_double)),
        ("population", POINTER(c_double)),
        ("commute_volume", POINTER(c_double)),
        ("weekday_c", POINTER(c_double)),
        ("log_area_c", POINTER(c_double)),
        ("initial_state", POINTER(c_double)),
        ("car_indptr", POINTER(c_int32)),
        ("car_indices", POINTER(c_int32)),
        ("car_values", POINTER(c_double)),
    ]


class SeirUpdateSpec(ctypes.Structure):
    """include/seir_b200.h: struct seir_update_spec."""

    _fields_ = [(n, c_int32) for n in ("kind", "target", "prev", "next", "mmax", "nmax", "dmax", "t0", "t1")]


class SeirSweepSpec(ctypes.Structure):
    """include/seir_b200.h: struct seir_sweep_spec."""

    _fields_ = [(n, c_int32) for n in ("num_leapfrog_steps", "num_event_time_updates", "dmax", "nmax", "mmax", "occult_nmax", "t0", "t1")] + [
        ("chain_offset", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("seed", ctypes.c_uint64)]


MMAX = 4  # columns of a proposal record [4][MMAX] (rows m, t, delta_t, x_star)


class NativeError(RuntimeError):
    pass


_LIB = None

# name -> (restype, argtypes); every symbol include/seir_b200.h declares
SIGNATURES = {
    "seir_abi_version": (c_int, []),
    "seir_last_error": (c_char_p, []),
    "seir_model_create": (c_int, [POINTER(SeirSpec), c_int, POINTER(c_void_p)]),
    "seir_model_destroy": (None, [c_void_p]),
    "seir_model_dims": (c_int, [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "seir_chains_create": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "seir_chains_destroy": (None, [c_void_p]),
    "seir_chains_bytes": (c_int64, [c_void_p]),
    "seir_compute_state": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "seir_ingest_events": (c_int, [c_void_p, c_void_p, c_void_p]),
    "seir_log_prob_cached": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "seir_log_prob": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "seir_log_prob_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "seir_log_prob_host_u16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "seir_last_h2d_bytes": (c_int64, [c_void_p]),
    "seir_log_prob_grad_cached": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "seir_run_stage": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "seir_prepare_theta": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "seir_update_step": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_hmc_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_hmc_draw": (c_int, [c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_propose": (c_int, [c_void_p, c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, c_void_p, c_void_p, c_void_p]),
    "seir_mcmc_sweep": (c_int, [c_void_p, c_void_p, ctypes.c_uint32] + [c_void_p] * 10),
    "seir_mcmc_burst": (c_int, [c_void_p, c_void_p, ctypes.c_uint32, ctypes.c_int32] + [c_void_p] * 10 + [ctypes.c_int32, c_void_p, c_void_p, c_void_p]),
    "seir_export_events": (c_int, [c_void_p, c_void_p, c_void_p]),
    "seir_export_events_u16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_simulate": (c_int, [c_void_p, c_int, ctypes.c_uint64, ctypes.c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_reproduction_number": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_pressure_components": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "seir_export_contraction": (c_int, [c_void_p, c_void_p, c_void_p]),
    "seir_chain_flags": (c_int, [c_void_p, c_void_p, c_void_p]),
    "seir_launch_count": (c_int64, []),
    "seir_sm_partition_info": (c_int, [c_int, c_void_p, c_void_p]),
}


def lib_path() -> str:
    """In-tree ``covid19uk_b200/lib/libseir_b200.so``; ``SEIR_B200_LIB`` names another build of the same ABI (A/B timing of
    kernel variants inside one GPU-box visit)."""
    return os.environ.get("SEIR_B200_LIB") or _build.lib_path()


def load():
    """Load the shared library once; raise loudly when it is absent (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise NativeError(
            f"{path} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m covid19uk_b200.build`."
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.seir_abi_version()
    if got != ABI_VERSION:
        raise NativeError(f"libseir_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild with --force")
    _LIB = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().seir_last_error()
        raise NativeError(f"seir_b200 error {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().seir_launch_count())


__all__ = [
    "SeirSpec", "NativeError", "load", "check", "lib_path", "launch_count", "SIGNATURES", "byref", "c_void_p",
    "THETA_CONSTRAINED", "THETA_UNCONSTRAINED", "PART_SEIR", "PART_PRIORS", "PART_ILDJ", "PART_JOINT",
]
