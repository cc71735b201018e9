"""Short driver for `ncu --set full` of the three hot kernels on the bench workload (config 4): one fp64 E+grad
evaluation with the static part (pair_fp64_kernel<POT,GRAD,FUSE>), one mixed E+grad (pair_diff_kernel<POT>) and one
mixed field-only pass (pair_diff_kernel<!POT>); a warm-up of each first."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
for _ in range(2):
    sysm.invalidate_static()
    e, g = sysm.energy_grad(D, precision="fp64")
    em, gm = sysm.energy_grad(D, precision="mixed")
    sysm.energy_grad(D, precision="mixed", want_energy=False)
torch.cuda.synchronize()
print("U_ind", float(e[0]), float(em[0]))
