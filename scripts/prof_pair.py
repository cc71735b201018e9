"""Short driver for `ncu --set full`: a few fp64 energy+gradient evaluations (and mixed field
passes) of the bench workload, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
for _ in range(3):
    e, g = sysm.energy_grad(D, precision="fp64")
for _ in range(3):
    sysm.energy_grad(D, precision="mixed", want_energy=False)
torch.cuda.synchronize()
print("U_ind", float(e[0]))
