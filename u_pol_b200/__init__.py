"""u_pol_b200 - B200-native (sm_100a CUDA) induced-dipole energy / gradient / minimiser path of
shehan807/U_pol, behind the reference's Python entry points.  See DESIGN.md."""
from .topology import DrudeSystem, build_system, ONE_4PI_EPS0  # noqa: F401

__all__ = ["DrudeSystem", "build_system", "ONE_4PI_EPS0"]
