"""TEST INFRASTRUCTURE - ctypes wrapper of oracle/uind_oracle.c (checker and CPU baseline).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libuind_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB):
            subprocess.run(["make", "-C", _HERE], check=True)
        h = C.CDLL(_LIB)
        h.upol_oracle_one_4pi_eps0.restype = C.c_double
        h.upol_oracle_max_threads.restype = C.c_int
        h.upol_oracle_energy_grad.restype = C.c_int
        h.upol_oracle_energy_grad.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 6 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        _lib = h
    return _lib


def max_threads():
    return int(lib().upol_oracle_max_threads())


def use_all_cores():
    """Undo an inherited OMP_NUM_THREADS=1 (torchrun sets it for its workers)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().upol_oracle_set_threads(int(n))
    return max_threads()


def energy_grad(r_core, d, q_core, q_shell, k, thole_u=None, row_lo=0, row_hi=None, want_grad=True):
    """r_core, d: (nmol,natoms,3); q_core, q_shell, k: (nmol,natoms); thole_u (nmol,natoms,natoms)|None.
    Returns (energy[4] = [U_ind, U_coul, U_coul_static, U_self] shares of the rows, grad (nmol,natoms,3))."""
    r = np.ascontiguousarray(r_core, dtype=np.float64)
    nmol, natoms = r.shape[0], r.shape[1]
    n = nmol * natoms
    dd = np.ascontiguousarray(np.asarray(d, dtype=np.float64).reshape(n, 3))
    qc, qs, kk = (np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(n)) for a in (q_core, q_shell, k))
    th = None
    if thole_u is not None and np.any(np.asarray(thole_u) != 0):
        th = np.ascontiguousarray(thole_u, dtype=np.float64)
    e = np.zeros(4)
    g = np.zeros((n, 3)) if want_grad else None
    rc = lib().upol_oracle_energy_grad(n, natoms, r.ctypes.data, dd.ctypes.data, qc.ctypes.data, qs.ctypes.data,
                                       kk.ctypes.data, th.ctypes.data if th is not None else None,
                                       row_lo, n if row_hi is None else row_hi, e.ctypes.data,
                                       g.ctypes.data if want_grad else None)
    if rc != 0:
        raise RuntimeError(f"upol_oracle_energy_grad failed: {rc}")
    return e, (g.reshape(nmol, natoms, 3) if want_grad else None)


_LIB_LD = os.path.join(_HERE, "libuind_oracle_ld.so")
_lib_ld = None


def uind_long_double(r_core, d, q_core, q_shell, k, thole_u=None):
    """U_ind from the 80-bit long-double arbiter (oracle/uind_oracle_ld.c); energy only."""
    global _lib_ld
    if _lib_ld is None:
        if not os.path.isfile(_LIB_LD):
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib_ld = C.CDLL(_LIB_LD)
        _lib_ld.upol_oracle_uind_long_double.restype = C.c_double
        _lib_ld.upol_oracle_uind_long_double.argtypes = [C.c_int64, C.c_int] + [C.c_void_p] * 6
    r = np.ascontiguousarray(r_core, dtype=np.float64)
    nmol, natoms = r.shape[0], r.shape[1]
    n = nmol * natoms
    dd = np.ascontiguousarray(np.asarray(d, dtype=np.float64).reshape(n, 3))
    qc, qs, kk = (np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(n)) for a in (q_core, q_shell, k))
    th = None
    if thole_u is not None and np.any(np.asarray(thole_u) != 0):
        th = np.ascontiguousarray(thole_u, dtype=np.float64)
    return float(_lib_ld.upol_oracle_uind_long_double(n, natoms, r.ctypes.data, dd.ctypes.data, qc.ctypes.data,
                                                      qs.ctypes.data, kk.ctypes.data,
                                                      th.ctypes.data if th is not None else None))
