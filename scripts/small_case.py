import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
s = synth.water_box(3000, seed=2)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
for _ in range(5): sysm.energy_grad(D)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): sysm.energy_grad(D)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("enqueue %.1f us/eval, total %.1f us/eval" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))
d, info = sysm.minimize(tol=1e-4)
t0 = time.perf_counter(); d, info = sysm.minimize(tol=1e-4); torch.cuda.synchronize(); print("minimise %.2f ms" % ((time.perf_counter() - t0) * 1e3), info["evaluations"])
