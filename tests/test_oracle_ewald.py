"""Pins the periodic (Ewald) oracle ``oracle/ewald_dense.py``: Madelung constants, independence of the
splitting parameter, and the open-boundary limit against the pinned open-boundary oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import ewald_dense, uind_dense
from oracle.uind_dense import COULOMB
from u_pol_b200 import synth


def test_madelung_nacl_and_cscl():
    a = 1.0
    fcc = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    X = np.concatenate([fcc, fcc + [.5, 0, 0]]) * a
    Q = np.array([1.0] * 4 + [-1.0] * 4)
    for kappa in (5.0, 7.0):
        e = float(ewald_energy_all(X, Q, [a] * 3, kappa))
        assert e / (4 * COULOMB / (a / 2)) == pytest.approx(-1.7475645946331822, abs=1e-12)
    X = np.array([[0, 0, 0], [.5, .5, .5]]) * a
    e = float(ewald_energy_all(X, np.array([1.0, -1.0]), [a] * 3, 6.0))
    assert e / (COULOMB / (a * math.sqrt(3) / 2)) == pytest.approx(-1.76267477307099, abs=1e-12)


def ewald_energy_all(X, Q, box, kappa):
    return ewald_dense.ewald_energy(X, Q, np.arange(len(Q)), box, kappa, images=2)


def _neutral_molecules(seed, nmol=6, na=3, L=2.0):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0, L, size=(nmol, 1, 3))
    X = (centres + rng.normal(scale=0.08, size=(nmol, na, 3))).reshape(-1, 3)
    q = rng.uniform(-1, 1, size=(nmol, na)); q -= q.mean(1, keepdims=True)
    return X, q.reshape(-1), np.repeat(np.arange(nmol), na)


def test_kappa_independence_with_exclusions_and_non_cubic_box():
    X, Q, mol = _neutral_molecules(1)
    box = [2.0, 2.3, 1.9]
    es = [float(ewald_dense.ewald_energy(X, Q, mol, box, kappa, images=2)) for kappa in (3.5, 4.5, 6.0)]
    assert abs(es[0] - es[1]) <= 1e-11 * abs(es[0]) and abs(es[0] - es[2]) <= 1e-11 * abs(es[0])
    # translation by a lattice vector and by an arbitrary vector changes nothing
    e_shift = float(ewald_dense.ewald_energy(X + [2.0, -2.3, 0.0], Q, mol, box, 4.5, images=2))
    assert abs(e_shift - es[1]) <= 1e-11 * abs(es[1])
    # minimum-image real space (images=0 -> home cell only ... here: images=1 vs 2) is complete at the default kappa
    kd = ewald_dense.default_kappa(box)
    e1 = float(ewald_dense.ewald_energy(X, Q, mol, box, kd, images=1))
    e2 = float(ewald_dense.ewald_energy(X, Q, mol, box, kd, images=2))
    assert abs(e1 - e2) <= 1e-13 * abs(e1) and abs(e1 - es[0]) <= 1e-11 * abs(es[0])


def test_open_boundary_limit_matches_the_pinned_oracle():
    s = synth.water_box(8, seed=5)
    D = synth.random_displacements(s, seed=6)
    R, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
    e_open, g_open = uind_dense.u_ind_and_grad(R, D, qis, qjs, qic, qjc, u, kk)
    diffs = []
    for L in (60.0, 120.0):
        box = [L] * 3
        e, g = ewald_dense.u_ind_pbc_and_grad(s.r_core + L / 2, D, s.q_core, s.q_shell, s.k, s.thole_u, box,
                                              ewald_dense.default_kappa(box))
        diffs.append(abs(e - e_open))
        assert np.abs(g - g_open).max() <= 1e-5 * np.abs(g_open).max()
    assert diffs[1] <= 2e-7 * abs(e_open)
    assert diffs[1] < diffs[0] / 4 or diffs[0] < 1e-9 * abs(e_open)      # image interactions decay ~ L^-3


def test_periodic_uind_kappa_independent_and_gradient_consistent():
    s = synth.imidazole_box(4, seed=7)
    D = synth.random_displacements(s, seed=8)
    L = float(np.ptp(s.r_core.reshape(-1, 3), axis=0).max() + 0.45)
    box = [L, L + 0.2, L + 0.1]
    k0 = ewald_dense.default_kappa(box)
    e0, g0 = ewald_dense.u_ind_pbc_and_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, box, k0)
    e1, g1 = ewald_dense.u_ind_pbc_and_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, box, 1.25 * k0, images=2)
    assert abs(e0 - e1) <= 1e-10 * abs(e0)
    assert np.abs(g0 - g1).max() <= 1e-10 * np.abs(g0).max()
    # central finite difference of the energy along a random direction
    rng = np.random.default_rng(0)
    v = np.where((s.q_shell != 0)[..., None], rng.normal(size=D.shape), 0.0)
    h = 1e-6
    ep = float(ewald_dense.u_ind_pbc(s.r_core, D + h * v, s.q_core, s.q_shell, s.k, s.thole_u, box, k0))
    em = float(ewald_dense.u_ind_pbc(s.r_core, D - h * v, s.q_core, s.q_shell, s.k, s.thole_u, box, k0))
    assert (ep - em) / (2 * h) == pytest.approx(float((g0 * v).sum()), rel=1e-6)
    # a periodic box differs from the isolated cluster
    R, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
    e_open = float(uind_dense.u_ind(R, D, qis, qjs, qic, qjc, u, kk))
    assert abs(e_open - e0) > 1e-4 * abs(e0)


def test_periodic_core_gradient_of_the_oracle():
    """dU_ind/dr_core under a box: no net force, agrees with central differences, kappa independent."""
    s = synth.water_box(6, seed=9)
    D = synth.random_displacements(s, seed=10)
    box = np.ptp(s.r_core.reshape(-1, 3), axis=0) + np.array([0.5, 0.6, 0.7])
    k0 = ewald_dense.default_kappa(box)
    args = (s.q_core, s.q_shell, s.k, s.thole_u, box)
    g = ewald_dense.u_ind_pbc_core_grad(s.r_core, D, *args, k0)
    assert np.abs(g.reshape(-1, 3).sum(0)).max() <= 1e-9 * np.abs(g).max()
    g2 = ewald_dense.u_ind_pbc_core_grad(s.r_core, D, *args, 1.2 * k0, images=2)
    assert np.abs(g - g2).max() <= 1e-9 * np.abs(g).max()
    rng = np.random.default_rng(1)
    v = rng.normal(size=s.r_core.shape)
    h = 1e-6
    ep = float(ewald_dense.u_ind_pbc(s.r_core + h * v, D, *args, k0))
    em = float(ewald_dense.u_ind_pbc(s.r_core - h * v, D, *args, k0))
    assert (ep - em) / (2 * h) == pytest.approx(float((g * v).sum()), rel=1e-5)
