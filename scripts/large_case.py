"""Beyond the BASELINE sizes: a 10^6-site water box (250 k Drude rows, 1.25 M columns) on one GPU.
The C oracle would need ~10 minutes for one evaluation, so parity is checked on sampled rows against a
direct numpy fp64 sum over ALL columns, plus fp64-vs-mixed agreement and the minimiser's own flags."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
from u_pol_b200.topology import ONE_4PI_EPS0 as K

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
s = synth.water_box(nmol, seed=9)
D = synth.random_displacements(s, seed=10)
sys_ = UindSystem.from_drude_system(s)
def timed(fn, n=2):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize(); return out, (time.perf_counter() - t0) / n
(e, g), t64 = timed(lambda: sys_.energy_grad(D))
(_, gm), tmx = timed(lambda: sys_.energy_grad(D, precision="mixed", want_energy=False))
g = g.cpu().numpy().reshape(-1, 3); gm = gm.cpu().numpy().reshape(-1, 3)
nsite = g.shape[0]
r = s.r_core.reshape(-1, 3); d = D.reshape(-1, 3); qc = s.q_core.reshape(-1); qs = s.q_shell.reshape(-1); k = s.k.reshape(-1)
na = s.r_core.shape[1]
shells = np.nonzero(qs)[0]
rng = np.random.default_rng(0)
rows = rng.choice(shells, 48, replace=False)
worst = 0.0
sh_pos = r[shells] - d[shells]
for i in rows:
    si = r[i] - d[i]
    m = i // na
    oc = np.ones(nsite, bool); oc[m * na:(m + 1) * na] = False
    os_ = (shells // na) != m
    xc = si - r[oc]; xs = si - sh_pos[os_]
    F = (qc[oc, None] * xc / (np.linalg.norm(xc, axis=1) ** 3)[:, None]).sum(0) + \
        (qs[shells][os_, None] * xs / (np.linalg.norm(xs, axis=1) ** 3)[:, None]).sum(0)
    # dU/dd_i = +K qs_i * sum_j q_j x/|x|^3 (x = s_i - x_j, s_i = r_i - d_i) + k d_i   (water: no Thole pairs)
    gi = K * qs[i] * F + k[i] * d[i]
    worst = max(worst, np.abs(gi - g[i]).max())
gmax = np.abs(g).max()
dmix = np.abs(gm - g).max() / gmax
pairs = sys_.interactions() + sys_.interactions(static_part=True)
print(f"nsite {nsite} rows {len(shells)}: fp64 step {t64*1e3:.1f} ms ({nsite*(nsite-1)/t64:.3e} site-pairs/s), mixed grad {tmx*1e3:.1f} ms; "
      f"sampled-row grad err {worst/gmax:.2e} (rel. to max |g|), mixed vs fp64 {dmix:.2e}", flush=True)
t0 = time.perf_counter(); dopt, info = sys_.minimize(tol=1e-4, precision="mixed"); torch.cuda.synchronize(); tm = time.perf_counter() - t0
print(f"minimise (mixed): {tm:.2f} s, {info['iterations']} iterations, {info['evaluations']} evaluations, flags {info['flags']}, "
      f"|g| {info['gnorm']:.2e}, U_ind {info['U_ind']:.10g}; memory {torch.cuda.max_memory_allocated()/2**20:.0f} MiB torch-side", flush=True)
assert worst / gmax < 1e-10 and dmix < 1e-5 and info["flags"] == 0
