"""GPU: batched small-system minimiser (BASELINE config 5; replaces wrapper.py:265-316's loop)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from u_pol_b200 import synth  # noqa: E402
from u_pol_b200.engine import UindSystem, batched_minimize  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
K = 138.93545764438198


def _dimers():
    z = np.load(os.path.join(GOLD, "dimers312.npz"))
    al, qs = z["alpha"], z["q_shell"]
    k = np.where(al == 0, 0.0, K * qs**2 / np.where(al == 0, 1.0, al))
    return z, k


def test_312_dimers_one_launch_vs_reference_csv():
    z, k = _dimers()
    r = z["r_core"].astype(np.float32).astype(np.float64)      # as the reference saw them (OpenMM CPU platform)
    d, e, it = batched_minimize(r, z["q_core"], z["q_shell"], k, z["thole_u"], tol=1e-6)
    e = e.cpu().numpy()
    assert not e[:, 6].any()                                    # no flags
    assert np.sqrt(e[:, 4]).max() <= 1e-6
    rel = np.abs(e[:, 0] - z["md_induction"]) / np.abs(z["md_induction"])
    assert np.median(rel) < 1e-6 and (rel < 1e-5).sum() >= 285 and rel.max() < 5e-4
    assert int(it.max()) < 100
    # same minima as the large-system path, system by system (sample)
    sys_ = UindSystem.from_compact(r[0], z["q_core"], z["q_shell"], k, z["thole_u"])
    for i in (0, 17, 150, 311):
        sys_.set_positions(r[i])
        _, info = sys_.minimize(tol=1e-6)
        # far dimers: U_ind ~ 1e-5 kJ/mol is 1e-8 of the static sum it is the remainder of (e[:, 5]);
        # fp64 round-off of that sum is the floor
        floor = 1e-14 * abs(e[i, 5])
        assert abs(info["U_ind"] - e[i, 0]) <= 1e-8 * abs(e[i, 0]) + floor
        ei, _ = sys_.energy_grad(d[i])
        assert abs(float(ei[0]) - e[i, 0]) <= 1e-10 * abs(e[i, 0]) + floor


def test_10k_perturbed_dimers():
    z, k = _dimers()
    r = synth.perturbed_dimers(z["r_core"], 10000, seed=5)
    d, e, it = batched_minimize(r, z["q_core"], z["q_shell"], k, z["thole_u"], tol=1e-4)
    e = e.cpu().numpy()
    assert e.shape == (10000, 8) and not e[:, 6].any()
    assert np.sqrt(e[:, 4]).max() <= 1e-4 and np.all(e[:, 0] < 0)
    sys_ = UindSystem.from_compact(r[0], z["q_core"], z["q_shell"], k, z["thole_u"])
    for i in (312, 5000, 9999):
        sys_.set_positions(r[i])
        _, info = sys_.minimize(tol=1e-6)
        assert abs(info["U_ind"] - e[i, 0]) <= 1e-8 * abs(e[i, 0]) + 1e-14 * abs(e[i, 5])


def test_10k_perturbed_dimers_sample_vs_newton_oracle():
    """A sample of the 10 000 perturbed dimers against the CPU oracle's exact (Newton) minimum of the same geometry -
    an external number for the batch, not the repo's own large-system engine."""
    from oracle import uind_dense                 # checker only

    z, k = _dimers()
    r = synth.perturbed_dimers(z["r_core"], 10000, seed=5)
    d, e, it = batched_minimize(r, z["q_core"], z["q_shell"], k, z["thole_u"], tol=1e-7)
    e = e.cpu().numpy()
    assert not e[:, 6].any()
    rng = np.random.default_rng(3)
    for i in [0, 311, 312, 9999] + list(rng.choice(10000, 8, replace=False)):
        R, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(r[i], z["q_core"], z["q_shell"], k, z["thole_u"])
        _, u_newton, gn = uind_dense.newton_minimum(R, np.zeros_like(r[i]), qis, qjs, qic, qjc, u, kk)
        assert gn <= 1e-8
        assert abs(e[i, 0] - u_newton) <= 1e-8 * abs(u_newton) + 1e-14 * abs(e[i, 5]), (i, e[i, 0], u_newton)


def test_warm_start_and_water():
    s = synth.water_box(2, seed=3)
    r = np.stack([s.r_core, s.r_core + 0.01])
    d, e, it = batched_minimize(r, s.q_core, s.q_shell, s.k, None, tol=1e-8)
    assert abs(float(e[0, 0]) - float(e[1, 0])) <= 1e-10 * abs(float(e[0, 0]))      # translation invariance
    d2, e2, it2 = batched_minimize(r, s.q_core, s.q_shell, s.k, None, d0=d, tol=1e-8)
    assert int(it2.max()) == 0 and float(e2[0, 0]) == pytest.approx(float(e[0, 0]), rel=1e-13)
