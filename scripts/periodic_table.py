"""Timings of the periodic (Ewald) path on liquid boxes in their own cell, next to the open-boundary
path on the same coordinates.  CUDA events, 3 warm-ups, median of 7."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

def timed(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

cases = [("water", synth.water_box, 15, 33.4), ("imidazole", synth.imidazole_box, 16, 7.0)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    cases.append(("imidazole", synth.imidazole_box, 27, 7.0))
for name, maker, nx, density in cases:
    nmol = nx ** 3
    s = maker(nmol, seed=2)
    a = (1.0 / density) ** (1.0 / 3.0)
    box = np.array([nx * a + 0.08] * 3)            # the lattice's own cell (+ margin: faces were not checked against images)
    D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
    sysm = UindSystem.from_drude_system(s)
    n = s.nmol * s.natoms
    def full():
        sysm.invalidate_static(); sysm.energy_grad(D)
    t_open_full = timed(full); t_open = timed(lambda: sysm.energy_grad(D))
    sysm.minimize(tol=1e-4, maxiter=2); torch.cuda.synchronize(); t0 = time.perf_counter()
    _, mo = sysm.minimize(tol=1e-4); torch.cuda.synchronize(); tm_open = time.perf_counter() - t0
    t_cg_open = timed(lambda: sysm.core_gradient(D))
    info = sysm.set_periodic(box)
    t_pbc_full = timed(full); t_pbc = timed(lambda: sysm.energy_grad(D))
    sysm.minimize(tol=1e-4, maxiter=2); torch.cuda.synchronize(); t0 = time.perf_counter()
    _, mp = sysm.minimize(tol=1e-4); torch.cuda.synchronize(); tm_pbc = time.perf_counter() - t0
    t_cg_pbc = timed(lambda: sysm.core_gradient(D))
    print(f"{name} {nmol} molecules: N = {n} sites, N_D = {int((s.q_shell != 0).sum())}, box {box[0]:.3f} nm, kappa {info['kappa']:.3f} /nm, {info['n_kvectors']} k-vectors")
    print(f"  open boundaries: E+grad {t_open_full:.3f} ms (static recomputed) / {t_open:.3f} ms (cached); minimise {tm_open*1e3:.1f} ms, {mo['iterations']} it, U = {mo['U_ind']:.6f}, flags {mo['flags']}; core gradient {t_cg_open:.3f} ms")
    print(f"  periodic (Ewald): E+grad {t_pbc_full:.3f} ms (static recomputed) / {t_pbc:.3f} ms (cached); minimise {tm_pbc*1e3:.1f} ms, {mp['iterations']} it, U = {mp['U_ind']:.6f}, flags {mp['flags']}; core gradient {t_cg_pbc:.3f} ms", flush=True)
    sysm.close()
