"""fp64 kernel variants on the 100 k-Drude box: full step (static recomputed), cached-static E+grad, potential only,
field only - CUDA-event kernel times.  Used with UPOL_B200_LIB for A/B runs of unroll factors / launch shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
def kt(fn, n=5):
    sysm.set_kernel_timing(True)
    for _ in range(2): fn()
    torch.cuda.synchronize(); sysm.kernel_time()
    for _ in range(n): fn()
    ms, k = sysm.kernel_time(); sysm.set_kernel_timing(False)
    return ms / k
def full():
    sysm.invalidate_static(); return sysm.energy_grad(D)
print(f"{os.path.basename(os.environ.get('UPOL_B200_LIB', 'default'))}: full {kt(full):.3f} ms  E+g cached {kt(lambda: sysm.energy_grad(D)):.3f} ms  "
      f"pot only {kt(lambda: sysm.energy_grad(D, want_grad=False)):.3f} ms  field only {kt(lambda: sysm.energy_grad(D, want_energy=False)):.3f} ms", flush=True)
