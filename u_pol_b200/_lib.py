"""ctypes binding of the C ABI (include/upol_b200.h) of libupol_b200.so.

The product has no CPU fallback: if the CUDA library is missing this module raises on first
use (``lib()``), and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UPOL_B200_LIB points at another build of the same library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("UPOL_B200_LIB") or os.path.join(_HERE, "lib", "libupol_b200.so")

E_UIND, E_INTER, E_INTRA, E_SELF, E_GNORM2, E_STATIC, E_SIZE = 0, 1, 2, 3, 4, 5, 8
FP64, MIXED = 0, 1
LBFGS, CG = 0, 1
FLAG_NOT_CONVERGED, FLAG_NONFINITE, FLAG_LINESEARCH, FLAG_COMM = 1, 2, 4, 8
COMM_ID_BYTES = 128
P2P_EXPORT_BYTES = 512


class MinOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("maxiter", C.c_int32), ("method", C.c_int32),
                ("precision", C.c_int32), ("history", C.c_int32), ("check_every", C.c_int32),
                ("reserved", C.c_int32)]


class MinResult(C.Structure):
    _fields_ = [("energy", C.c_double * E_SIZE), ("gnorm", C.c_double), ("iterations", C.c_int32),
                ("evaluations", C.c_int32), ("flags", C.c_int32), ("reserved", C.c_int32)]


class UpolError(RuntimeError):
    pass


_P = C.c_void_p
_I64 = C.c_int64
_SIGNATURES = {
    "upol_error_string": (C.c_char_p, [C.c_int]),
    "upol_version": (C.c_int, []),
    "upol_one_4pi_eps0": (C.c_double, []),
    "upol_device_count": (C.c_int, []),
    "upol_system_create": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _P, _P, _P, _P, _P, _I64, _P, _P, _P]),
    "upol_system_destroy": (C.c_int, [_P]),
    "upol_system_set_positions": (C.c_int, [_P, _P, _P]),
    "upol_system_set_positions_dev": (C.c_int, [_P, _P, _P]),
    "upol_system_nsite": (_I64, [_P]),
    "upol_system_ndrude": (_I64, [_P]),
    "upol_system_interactions": (_I64, [_P, C.c_int]),
    "upol_system_set_row_range": (C.c_int, [_P, _I64, _I64]),
    "upol_uind_energy_grad_dev": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "upol_uind_energy_grad_host": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "upol_uind_energy_grad_host_sharded": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.POINTER(_I64), C.POINTER(_I64)]),
    "upol_uind_core_gradient_dev": (C.c_int, [_P, _P, _P, _P]),
    "upol_coulomb_static_dev": (C.c_int, [_P, _P, _P]),
    "upol_uself_dev": (C.c_int, [_P, _P, _I64, _P, _P]),
    "upol_make_sij_dev": (C.c_int, [_P, _P, C.c_int, _I64, _P, _P]),
    "upol_min_options_default": (None, [C.POINTER(MinOptions)]),
    "upol_minimize_dev": (C.c_int, [_P, _P, C.POINTER(MinOptions), C.POINTER(MinResult), _P]),
    "upol_minimize_host": (C.c_int, [_P, _P, C.POINTER(MinOptions), C.POINTER(MinResult)]),
    "upol_comm_unique_id": (C.c_int, [_P]),
    "upol_system_attach_comm": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "upol_system_p2p_export": (C.c_int, [_P, _P]),
    "upol_system_p2p_attach": (C.c_int, [_P, _P]),
    "upol_system_p2p_disable": (C.c_int, [_P]),
    "upol_batched_minimize_dev": (C.c_int, [C.c_int, _I64, _I64, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P,
                                            C.c_double, C.c_int32, _P, _P, _P]),
    "upol_system_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "upol_system_launch_count": (_I64, [_P]),
    "upol_system_kernel_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "upol_debug_check_guards": (C.c_int, []),
    "upol_debug_guard_selftest": (C.c_int, []),
    "upol_measure_fma_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "upol_system_set_periodic": (C.c_int, [_P, _P, C.c_double, C.c_int32]),
    "upol_system_periodic_info": (C.c_int, [_P, _P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load libupol_b200.so (built in-tree by ``__graft_entry__.build()``); fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise UpolError(
                f"{LIB_PATH} not found: build the CUDA library first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check_min_flags(flags, what="minimisation"):
    """Minimiser flags (UPOL_FLAG_*): a lost peer or a non-finite gradient means the returned displacements
    are not a minimum - raise; an unconverged run or an exhausted line search is reported as a warning."""
    import warnings

    if flags & FLAG_COMM:
        raise UpolError(f"{what}: a peer rank did not deliver within the time-out (UPOL_FLAG_COMM); result invalid")
    if flags & FLAG_NONFINITE:
        raise UpolError(f"{what}: non-finite gradient (UPOL_FLAG_NONFINITE) - polarisation catastrophe or bad input")
    if flags & FLAG_NOT_CONVERGED:
        warnings.warn(f"{what}: not converged to the requested tolerance (UPOL_FLAG_NOT_CONVERGED)", RuntimeWarning)
    elif flags & FLAG_LINESEARCH:
        warnings.warn(f"{what}: a line search hit its step limit (UPOL_FLAG_LINESEARCH)", RuntimeWarning)
    return flags


def check(rc, what=""):
    if rc < 0:
        msg = lib().upol_error_string(rc).decode()
        raise UpolError(f"{what or 'upol call'} failed ({rc}): {msg}")
    return rc
