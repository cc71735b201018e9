"""Smallest case that touches every kernel (for compute-sanitizer, one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth, polarization as pol, util
from u_pol_b200.engine import UindSystem, batched_minimize

s = synth.imidazole_box(30, seed=1)
D = synth.random_displacements(s, seed=2)
sysm = UindSystem.from_drude_system(s)
e, g = sysm.energy_grad(D)
em, gm = sysm.energy_grad(D, precision="mixed")
eh, gh = sysm.energy_grad_host(D)
es, gs, rng = sysm.energy_grad_host_sharded(np.asarray(D).reshape(-1), r_core=np.ascontiguousarray(s.r_core.reshape(-1)))
sysm.energy_grad(D, want_energy=False); sysm.energy_grad(D, precision="mixed", want_energy=False)
cg = sysm.core_gradient(D)
sysm.set_row_range(3, 17); sysm.energy_grad(D); sysm.set_row_range(0, 30)
d, info = sysm.minimize(tol=1e-5)
d2, info2 = sysm.minimize(tol=1e-5, precision="mixed", method="cg")
w = synth.water_box(27, seed=3)
Rij, Dij = util.get_Rij_Dij(w); Qic, Qis, Qjc, Qjs = util.get_QiQj(w); k, u = util.get_pol_params(w)
pol.set_constants()
print(float(pol.Uind(Rij, Dij, Qis, Qjs, Qic, Qjc, u, k)), float(pol.Ucoul(Rij, Dij, Qis, Qjs, Qic, Qjc, u)),
      float(pol.Ucoul_static(Rij, Qis, Qjs, Qic, Qjc)), float(pol.Uself(Dij, k)), float(pol.make_Sij(Rij, 1.5).sum()))
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dimers312.npz"))
al, qs = z["alpha"], z["q_shell"]
kk = np.where(al == 0, 0.0, 138.93545764438198 * qs**2 / np.where(al == 0, 1.0, al))
dd, ee, it = batched_minimize(z["r_core"][:40], z["q_core"], qs, kk, z["thole_u"], tol=1e-6)
torch.cuda.synchronize()
print("ok", float(e[0]), info["U_ind"], info2["U_ind"], float(ee[0, 0]), int(it.max()))
