"""GPU: randomly generated topologies (the survey's 'hypothesis-generated small boxes'): random
molecule counts and sizes, random subsets of polarisable sites (including molecules without any),
random screened pairs, random charges.  Energy, gradient and minimum against the dense oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import uind_dense  # noqa: E402  (checker only)
from u_pol_b200.engine import UindSystem  # noqa: E402
from u_pol_b200.topology import ONE_4PI_EPS0  # noqa: E402


def _random_system(seed):
    rng = np.random.default_rng(seed)
    nmol, natoms = int(rng.integers(2, 9)), int(rng.integers(2, 8))
    # molecule geometry: sites on a small random cluster, molecules on a loose random lattice
    geom = rng.normal(scale=0.08, size=(natoms, 3))
    for a in range(1, natoms):                       # keep intra-molecular sites >= 0.09 nm apart
        while np.linalg.norm(geom[a] - geom[:a], axis=1).min() < 0.09:
            geom[a] = rng.normal(scale=0.12, size=3)
    cells = rng.permutation(27)[:nmol]
    centres = np.stack([cells // 9, (cells // 3) % 3, cells % 3], 1) * 0.75 + rng.normal(scale=0.03, size=(nmol, 3))
    r = centres[:, None, :] + geom[None, :, :] + rng.normal(scale=0.004, size=(nmol, natoms, 3))
    pol = rng.random(natoms) < 0.6
    if seed % 4 == 0:
        pol[:] = False                                # no polarisable site at all in one case out of four?
        pol[rng.integers(natoms)] = True
    qs = np.where(pol, -rng.uniform(0.6, 1.6, natoms), 0.0)
    qc = rng.uniform(-0.5, 0.5, natoms) - qs * pol    # core carries the Drude's counter-charge
    alpha = np.where(pol, rng.uniform(0.0008, 0.002, natoms), 0.0)
    th = np.zeros((natoms, natoms))
    for a in range(natoms):
        for b in range(a + 1, natoms):
            if pol[a] and pol[b] and rng.random() < 0.7:
                th[a, b] = th[b, a] = 2.0 / (alpha[a] * alpha[b]) ** (1.0 / 6.0)
    tile = lambda v: np.broadcast_to(v, (nmol,) + v.shape).copy()
    QC, QS, AL, TH = tile(qc), tile(qs), tile(alpha), tile(th)
    if nmol > 2:                                      # one molecule without any Drude
        QC[-1] += QS[-1]; QS[-1] = 0.0; AL[-1] = 0.0; TH[-1] = 0.0
    K = np.where(AL == 0, 0.0, ONE_4PI_EPS0 * QS**2 / np.where(AL == 0, 1.0, AL))
    D = np.where((QS != 0)[..., None], rng.normal(scale=0.002, size=r.shape), 0.0)
    return r, QC, QS, K, TH, D


@pytest.mark.parametrize("seed", range(12))
def test_random_topology(seed):
    r, qc, qs, k, th, D = _random_system(seed)
    Rij, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(r, qc, qs, k, th)
    e_ref, g_ref = uind_dense.u_ind_and_grad(Rij, D, qis, qjs, qic, qjc, u, kk)
    sys_ = UindSystem.from_compact(r, qc, qs, k, th)
    e, g = sys_.energy_grad(D)
    scale = max(abs(e_ref), 1e-3 * float(np.abs(k * (D**2).sum(-1)).sum()), 1e-6)
    assert abs(float(e[0]) - e_ref) <= 1e-10 * scale + 1e-12 * abs(float(e[5]))
    g = g.cpu().numpy().reshape(g_ref.shape)
    assert np.abs(g - g_ref).max() <= 1e-10 * max(np.abs(g_ref).max(), 1e-6)
    gm = sys_.energy_grad(D, precision="mixed", want_energy=False)[1].cpu().numpy().reshape(g_ref.shape)
    assert np.abs(gm - g_ref).max() <= 1e-5 * max(np.abs(g_ref).max(), 1e-6)
    # minimum: Newton on the oracle vs the device minimiser
    _, e_min, gn = uind_dense.newton_minimum(Rij, np.zeros_like(r), qis, qjs, qic, qjc, u, kk)
    d, info = sys_.minimize(tol=1e-6)
    if gn < 1e-7 and info["flags"] == 0:             # both converged (no polarisation catastrophe in the draw)
        assert abs(info["U_ind"] - e_min) <= 1e-8 * abs(e_min) + 1e-12 * abs(float(e[5]))
    cg = sys_.core_gradient(D).cpu().numpy()
    assert np.abs(cg.sum(0)).max() <= 1e-8 * max(np.abs(cg).max(), 1e-6) * len(cg)
