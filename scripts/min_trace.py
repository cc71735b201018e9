import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench
from u_pol_b200.engine import UindSystem
s, d, name = bench._config_workload(3)
sysm = UindSystem.from_drude_system(s)
_, info = sysm.minimize(tol=1e-4, check_every=1)
print(info["iterations"], info["evaluations"])
s = bench._box("water", 3000, 2)
sysm = UindSystem.from_drude_system(s)
_, info = sysm.minimize(tol=1e-4, check_every=1)
print(info["iterations"], info["evaluations"])
