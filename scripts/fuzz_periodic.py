"""Fuzz of the periodic (Ewald) path: random topologies (tests/test_gpu_random_topologies.py's
generator) in random orthorhombic boxes with random kappa, against oracle/ewald_dense.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_random_topologies import _random_system
from oracle import ewald_dense
from u_pol_b200.engine import UindSystem

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
worst = [0.0, 0.0, 0.0, 0.0]
for seed in range(lo, hi):
    r, qc, qs, k, th, D = _random_system(seed)
    rng = np.random.default_rng(50_000 + seed)
    ext = np.ptp(r.reshape(-1, 3), axis=0)
    box = ext + rng.uniform(0.3, 1.5, size=3)
    r = r + rng.uniform(-3, 3, size=3)                                   # anywhere relative to the cell
    shift = rng.integers(-2, 3, size=(r.shape[0], 1, 3)) * box           # molecules in different image cells
    kappa = rng.uniform(13.0, 17.0) / box.min()
    sys_ = UindSystem.from_compact(r + shift, qc, qs, k, th)
    sys_.set_periodic(box, kappa=kappa)
    e, g = sys_.energy_grad(D)
    e_ref, g_ref = ewald_dense.u_ind_pbc_and_grad(r, D, qc, qs, k, th, box, kappa, images=1)
    g = g.cpu().numpy().reshape(g_ref.shape)
    de = abs(float(e[0]) - e_ref) / (abs(e_ref) + 1e-4 * abs(float(e[5])) + 1e-12)
    dg = np.abs(g - g_ref).max() / max(np.abs(g_ref).max(), 1e-6)
    cg = sys_.core_gradient(D).cpu().numpy()
    cg_ref = ewald_dense.u_ind_pbc_core_grad(r, D, qc, qs, k, th, box, kappa).reshape(cg.shape)
    dc = np.abs(cg - cg_ref).max() / max(np.abs(cg_ref).max(), 1e-6)
    d, info = sys_.minimize(tol=1e-6)
    dmin = 0.0
    if info["flags"] == 0:
        e_at, g_at = ewald_dense.u_ind_pbc_and_grad(r, d.cpu().numpy().reshape(r.shape), qc, qs, k, th, box, kappa)
        dmin = max(abs(info["U_ind"] - e_at) / (abs(e_at) + 1e-4 * abs(float(e[5])) + 1e-12), float(np.linalg.norm(g_at)) * 1e-4)
    worst = [max(a, b) for a, b in zip(worst, (de, dg, dmin, dc))]
    if de > 1e-10 or dg > 1e-10 or dmin > 1e-8 or dc > 1e-9:
        bad += 1
        print(f"seed {seed}: dE {de:.2e} dg {dg:.2e} dcore {dc:.2e} dmin {dmin:.2e} flags {info['flags']} box {box} kappa {kappa:.3f}", flush=True)
    sys_.close()
print(f"seeds {lo}..{hi - 1}: {bad} failing; worst dE {worst[0]:.2e} dg {worst[1]:.2e} minimum/stationarity {worst[2]:.2e} core gradient {worst[3]:.2e}")
sys.exit(min(bad, 100))
