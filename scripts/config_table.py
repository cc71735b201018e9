"""Timings of the BASELINE.json configs 2, 3 and 5 on one GPU (config 4 is bench.py's workload).
Writes a small table to stdout; CUDA events, 3 warm-ups, median of 10."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem, batched_minimize

def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

print("config 2: SWM4-NDP-style water, 3000 molecules, N=12000 sites, N_D=3000, single E+grad")
s = synth.water_box(3000, seed=2)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
n = s.nmol * s.natoms
def full():
    sysm.invalidate_static(); sysm.energy_grad(D)
ms = timed(full)
print(f"  fp64 full E+grad (static recomputed): {ms*1e3:.1f} us  -> {n*(n-1)/ms/1e6:.2f} G site-pairs/s, {(sysm.interactions()+sysm.interactions(True))/ms/1e6:.1f} G interactions/s")
ms = timed(lambda: sysm.energy_grad(D))
print(f"  fp64 E+grad (static cached):          {ms*1e3:.1f} us")
ms = timed(lambda: sysm.energy_grad(D, precision='mixed', want_energy=False))
print(f"  mixed field only:                     {ms*1e3:.1f} us")
for prec in ("fp64", "mixed"):
    sysm.minimize(tol=1e-4, precision=prec, maxiter=2)           # warm: workspace allocation, module load
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, info = sysm.minimize(tol=1e-4, precision=prec); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"  minimise {prec} tol 1e-4: {t*1e3:.1f} ms, {info['iterations']} it / {info['evaluations']} evals, U = {info['U_ind']:.6f}, |g| = {info['gnorm']:.2e}")

print("config 3: imidazole-parameter liquid, 4000 molecules, N=36000, N_D=20000, full minimisation, 1 GPU")
s = synth.imidazole_box(4000, seed=3)
sysm = UindSystem.from_drude_system(s)
for prec in ("fp64", "mixed"):
    for tol in (1e-4, 1e-6):
        sysm.minimize(tol=tol, precision=prec, maxiter=2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d, info = sysm.minimize(tol=tol, precision=prec); torch.cuda.synchronize(); t = time.perf_counter() - t0
        print(f"  {prec:5s} tol {tol:.0e}: {t*1e3:.1f} ms, {info['iterations']} it / {info['evaluations']} evals, U = {info['U_ind']:.6f}, |g| = {info['gnorm']:.2e}, flags {info['flags']}")

print("config 5: 10000 imidazole dimers (312 reference geometries + rigid-body perturbations), one launch")
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dimers312.npz"))
al, qs = z["alpha"], z["q_shell"]
k = np.where(al == 0, 0.0, 138.93545764438198 * qs**2 / np.where(al == 0, 1.0, al))
r = torch.as_tensor(synth.perturbed_dimers(z["r_core"], 10000, seed=5)).cuda()
for tol in (1e-4, 1e-6):
    batched_minimize(r, z["q_core"], qs, k, z["thole_u"], tol=tol)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, e, it = batched_minimize(r, z["q_core"], qs, k, z["thole_u"], tol=tol); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(f"  tol {tol:.0e}: {t*1e3:.2f} ms total -> {10000/t:.0f} systems/s; iterations mean {float(it.float().mean()):.1f} max {int(it.max())}; flags any {bool(e[:,6].any())}")
