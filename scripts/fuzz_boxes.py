"""Fuzz at medium size: random water / imidazole boxes (30-900 molecules, random seeds) - energy and
gradient against the fp64 C oracle (energy bar 5e-10: the oracle itself carries ~1e-10), minimiser convergence (fp64, mixed, CG) and agreement of the minima."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import c_oracle
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(2024)
bad = 0
worst = dict(dE=0.0, dg=0.0, dmix=0.0, dmin=0.0, it=0, dEm=0.0)
for case in range(n_cases):
    maker = synth.water_box if case % 2 == 0 else synth.imidazole_box
    nmol = int(rng.integers(30, 900))
    seed = int(rng.integers(1, 10**6))
    s = maker(nmol, seed=seed)
    D = synth.random_displacements(s, seed=seed + 1, sigma=float(rng.uniform(0.0005, 0.004)))
    e_ref, g_ref = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
    sys_ = UindSystem.from_drude_system(s)
    e, g = sys_.energy_grad(D)
    g = g.cpu().numpy().reshape(g_ref.shape)
    gm = sys_.energy_grad(D, precision="mixed", want_energy=False)[1].cpu().numpy().reshape(g_ref.shape)
    em, gme = sys_.energy_grad(D, precision="mixed")             # energy AND gradient from the fp32 difference-form pass
    dEm = abs(float(em[0]) - e_ref[0]) / abs(e_ref[0])
    dgme = np.abs(gme.cpu().numpy().reshape(g_ref.shape) - g_ref).max() / np.abs(g_ref).max()
    dE = abs(float(e[0]) - e_ref[0]) / abs(e_ref[0])
    dg = np.abs(g - g_ref).max() / np.abs(g_ref).max()
    dmix = np.abs(gm - g_ref).max() / np.abs(g_ref).max()
    res = {}
    for name, kw in (("fp64", {}), ("mixed", {"precision": "mixed"}), ("cg", {"method": "cg"})):
        _, info = sys_.minimize(tol=1e-4, **kw)
        res[name] = info
    dmin = max(abs(res[k]["U_ind"] - res["fp64"]["U_ind"]) / abs(res["fp64"]["U_ind"]) for k in res)
    flags = [res[k]["flags"] for k in res]
    its = [res[k]["iterations"] for k in res]
    worst["dE"] = max(worst["dE"], dE); worst["dg"] = max(worst["dg"], dg); worst["dmix"] = max(worst["dmix"], dmix)
    worst["dmin"] = max(worst["dmin"], dmin); worst["it"] = max(worst["it"], max(its[:2]))
    worst["dEm"] = max(worst["dEm"], dEm); worst["dmix"] = max(worst["dmix"], dgme)
    if dE > 5e-10 or dg > 1e-10 or dmix > 1e-5 or dmin > 1e-8 or any(flags) or dEm > 1e-5 or dgme > 1e-5:
        bad += 1
        print(f"case {case} {maker.__name__} nmol {nmol} seed {seed}: dE {dE:.1e} dg {dg:.1e} dmix {dmix:.1e} dEmixed {dEm:.1e} dmin {dmin:.1e} flags {flags} its {its}", flush=True)
    sys_.close()
print(f"{n_cases} boxes: {bad} failing; worst dE {worst['dE']:.2e} dg {worst['dg']:.2e} dmixed {worst['dmix']:.2e} dE_mixed {worst['dEm']:.2e} dmin {worst['dmin']:.2e} max L-BFGS iterations {worst['it']}")
sys.exit(min(bad, 100))
