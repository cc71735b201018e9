"""Fuzz: many random topologies (tests/test_gpu_random_topologies.py's generator) on the GPU against
the dense oracle.  Prints failures; exit code = number of failing seeds."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_random_topologies import _random_system
from oracle import uind_dense
from u_pol_b200.engine import UindSystem

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
worst = [0.0, 0.0, 0.0, 0.0, 0.0]
for seed in range(lo, hi):
    r, qc, qs, k, th, D = _random_system(seed)
    Rij, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(r, qc, qs, k, th)
    e_ref, g_ref = uind_dense.u_ind_and_grad(Rij, D, qis, qjs, qic, qjc, u, kk)
    sys_ = UindSystem.from_compact(r, qc, qs, k, th)
    e, g = sys_.energy_grad(D)
    g = g.cpu().numpy().reshape(g_ref.shape)
    gm = sys_.energy_grad(D, precision="mixed", want_energy=False)[1].cpu().numpy().reshape(g_ref.shape)
    em, gme = sys_.energy_grad(D, precision="mixed")             # mixed energy and gradient from one fp32 pass
    scale = max(abs(e_ref), 1e-3 * float(np.abs(k * (D**2).sum(-1)).sum()), 1e-6)
    de = abs(float(e[0]) - e_ref) / (scale + 1e-2 * abs(float(e[5])))
    dg = np.abs(g - g_ref).max() / max(np.abs(g_ref).max(), 1e-6)
    dm = np.abs(gm - g_ref).max() / max(np.abs(g_ref).max(), 1e-6)
    dm = max(dm, np.abs(gme.cpu().numpy().reshape(g_ref.shape) - g_ref).max() / max(np.abs(g_ref).max(), 1e-6))
    dem = abs(float(em[0]) - e_ref) / (scale + 1e-2 * abs(float(e[5])))
    _, e_min, gn = uind_dense.newton_minimum(Rij, np.zeros_like(r), qis, qjs, qic, qjc, u, kk)
    d, info = sys_.minimize(tol=1e-6)
    dmin = 0.0
    if gn < 1e-7 and info["flags"] == 0:
        dmin = abs(info["U_ind"] - e_min) / (abs(e_min) + 1e-4 * abs(float(e[5])) + 1e-30)
    worst = [max(a, b) for a, b in zip(worst, (de, dg, dm, dmin, dem))]
    if de > 1e-10 or dg > 1e-10 or dm > 1e-5 or dem > 1e-5 or dmin > 1e-8 or (gn < 1e-7 and info["flags"] not in (0,)):
        bad += 1
        print(f"seed {seed}: dE {de:.2e} dg {dg:.2e} dmixed {dm:.2e} dE_mixed {dem:.2e} dmin {dmin:.2e} flags {info['flags']} newton_gn {gn:.1e} it {info['iterations']}", flush=True)
    sys_.close()
print(f"seeds {lo}..{hi - 1}: {bad} failing; worst dE {worst[0]:.2e} dg {worst[1]:.2e} dmixed {worst[2]:.2e} dE_mixed {worst[4]:.2e} dmin {worst[3]:.2e}")
sys.exit(min(bad, 100))
