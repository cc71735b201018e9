"""Host-side array construction with the reference's names (U_pol/util.py:37-279), taking a
:class:`u_pol_b200.topology.DrudeSystem` (built from PDB + force-field XML without OpenMM)
where the reference takes an OpenMM ``Simulation``.  Outputs have the reference's dense
broadcast layout, as numpy fp64 arrays; they feed ``u_pol_b200.polarization`` directly.
These are O(N^2) in memory exactly like the reference's - use ``UindSystem`` for large N.
"""
from __future__ import annotations

import numpy as np

from .topology import ONE_4PI_EPS0, DrudeSystem, build_system  # noqa: F401


def set_constants():
    """util.py:28-34."""
    return ONE_4PI_EPS0


def get_Dij(r_core, r_shell):
    """util.py:37-42: d = r_core - r_shell, zero where the shell position is the origin."""
    r_core, r_shell = np.asarray(r_core, dtype=np.float64), np.asarray(r_shell, dtype=np.float64)
    mask = np.linalg.norm(r_shell, axis=-1) > 0.0
    return np.where(mask[..., None], r_core - r_shell, 0.0)


def get_Rij_Dij(system: DrudeSystem):
    """util.py:45-84: Rij[i,j,a,b] = r[j,b] - r[i,a]; Dij starts at zero (shell on core)."""
    r = np.asarray(system.r_core, dtype=np.float64)
    Rij = r[None, :, None, :, :] - r[:, None, :, None, :]
    return Rij, get_Dij(r, r.copy())


def get_QiQj(system: DrudeSystem):
    """util.py:87-147: (Qi_core, Qi_shell, Qj_core, Qj_shell) broadcast views."""
    qc, qs = np.asarray(system.q_core, dtype=np.float64), np.asarray(system.q_shell, dtype=np.float64)
    return qc[:, None, :, None], qs[:, None, :, None], qc[None, :, None, :], qs[None, :, None, :]


def get_pol_params(system: DrudeSystem):
    """util.py:150-279: (k, u_scale); u_scale is the Python float 0.0 without screened pairs."""
    k = system.k
    if system.has_screened_pairs:
        u_scale = system.thole_u[None, ...] * np.eye(system.nmol)[:, :, None, None]
    else:
        u_scale = 0.0
    return k, u_scale
