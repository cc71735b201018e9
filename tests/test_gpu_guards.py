"""GPU: out-of-bounds stores next to the library's device buffers (compute-sanitizer is closed on the GPU pool, so
the library can allocate every buffer with guard zones, UPOL_GUARD=1, and check them afterwards - csrc/guard.cu).
The all-kernel smallest case (scripts/sanitize_case.py) plus shapes that stress the padding rules run in a
subprocess with the guards on; no guard zone may be touched."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

CASE = r'''
import os, sys, runpy
sys.path.insert(0, ROOT)
import numpy as np, torch
from u_pol_b200 import _lib as L, synth
from u_pol_b200.engine import UindSystem
assert L.lib().upol_debug_check_guards() == 0
assert L.lib().upol_debug_guard_selftest() == 1, "the guard mechanism does not see a one-byte overrun"
runpy.run_path(os.path.join(ROOT, "scripts", "sanitize_case.py"), run_name="__main__")
assert L.lib().upol_debug_check_guards() == 0, "guard zone damaged by the all-kernel case"
# sizes around the tile / row-block / split boundaries, water (1 Drude per molecule) and imidazole (5 of 9)
keep = []
for maker, nmol in ((synth.water_box, 1), (synth.water_box, 127), (synth.water_box, 129), (synth.imidazole_box, 26),
                    (synth.imidazole_box, 103), (synth.water_box, 513), (synth.imidazole_box, 410)):
    s = maker(nmol, seed=nmol)
    D = synth.random_displacements(s, seed=1)
    sysm = UindSystem.from_drude_system(s)
    for prec in ("fp64", "mixed"):
        sysm.energy_grad(D, precision=prec)
        sysm.energy_grad(D, precision=prec, want_energy=False)
        sysm.minimize(tol=1e-5, precision=prec, maxiter=30)
    sysm.set_row_range(0, max(1, nmol // 3)); sysm.energy_grad(D); sysm.minimize(tol=1e-5, maxiter=5)
    sysm.core_gradient(D)
    if nmol >= 26:
        L_box = (s.nmol * s.natoms / 60.0) ** (1 / 3) + 1.5
        sysm.set_row_range(0, nmol)
        sysm.set_periodic([max(L_box, 3.0)] * 3); sysm.energy_grad(D); sysm.minimize(tol=1e-4, maxiter=5); sysm.set_periodic(None)
    keep.append(sysm)
    bad = L.lib().upol_debug_check_guards()
    assert bad == 0, (maker.__name__, nmol, bad)
torch.cuda.synchronize()
print("guards clean:", L.lib().upol_debug_check_guards())
'''


def test_no_store_outside_any_device_buffer():
    env = dict(os.environ, UPOL_GUARD="1")
    out = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + CASE], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert "guards clean: 0" in out.stdout


def test_guard_mode_is_off_by_default():
    from u_pol_b200 import _lib as L

    if os.environ.get("UPOL_GUARD"):
        pytest.skip("guard mode forced by the environment")
    assert L.lib().upol_debug_check_guards() == -1
