"""Short driver for ncu: a few periodic (Ewald) evaluations of a 4096-imidazole box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
s = synth.imidazole_box(4096, seed=2)
box = np.array([16 * (1.0 / 7.0) ** (1.0 / 3.0) + 0.08] * 3)
D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
sysm = UindSystem.from_drude_system(s)
sysm.set_periodic(box)
for _ in range(3):
    sysm.invalidate_static()
    e, g = sysm.energy_grad(D)
torch.cuda.synchronize()
print("U_ind", float(e[0]))
