"""Short driver for `ncu --set full`: one mixed energy+gradient pass and one mixed field-only pass
(pair_diff_kernel<POT=true/false>) of the bench workload, nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
e, g = sysm.energy_grad(D, precision="mixed")
sysm.energy_grad(D, precision="mixed", want_energy=False)
torch.cuda.synchronize()
print("U_ind", float(e[0]))
