"""Config 2 (3000 waters): a few evaluations of each kind and one short minimisation, for a launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
s = synth.water_box(3000, seed=2)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
for _ in range(2):
    sysm.invalidate_static(); sysm.energy_grad(D)
    sysm.energy_grad(D, precision="mixed")
    sysm.energy_grad(D, precision="mixed", want_energy=False)
    sysm.energy_grad(D, want_energy=False)
torch.cuda.synchronize()
d, info = sysm.minimize(tol=1e-4, maxiter=3)
torch.cuda.synchronize()
print(info["iterations"], info["U_ind"])
