"""Host enqueue time vs device time of one evaluation at config-2 size (3000 waters)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
s = synth.water_box(3000, seed=2)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
def run(label, fn, n=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label}: enqueue {(t1 - t0) / n * 1e6:.1f} us/eval, total {(t2 - t0) / n * 1e6:.1f} us/eval", flush=True)
run("fp64 E+grad", lambda: sysm.energy_grad(D))
run("fp64 grad only", lambda: sysm.energy_grad(D, want_energy=False))
run("fp64 E only", lambda: sysm.energy_grad(D, want_grad=False))
run("mixed grad only", lambda: sysm.energy_grad(D, precision="mixed", want_energy=False))
def full():
    sysm.invalidate_static(); sysm.energy_grad(D)
run("fp64 E+grad, static recomputed", full)
run("fp64 E+grad again", lambda: sysm.energy_grad(D))
d, info = sysm.minimize(tol=1e-4)
for _ in range(3):
    t0 = time.perf_counter(); d, info = sysm.minimize(tol=1e-4); torch.cuda.synchronize()
    print("minimise %.2f ms" % ((time.perf_counter() - t0) * 1e3), info["evaluations"], flush=True)
