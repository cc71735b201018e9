"""CPU: the C restatement (oracle/uind_oracle.c, the CPU-baseline code) against the dense torch oracle."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import uind_dense as O
from u_pol_b200 import synth


@pytest.mark.parametrize("maker,nmol,seed", [(synth.water_box, 27, 1), (synth.imidazole_box, 20, 2), (synth.imidazole_box, 2, 3)])
def test_c_oracle_vs_dense(maker, nmol, seed):
    s = maker(nmol, seed=seed)
    D = np.random.default_rng(seed).normal(scale=0.002, size=s.r_core.shape)      # all sites displaced
    R, qis, qjs, qic, qjc, u, k = O.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
    e_ref, g_ref = O.u_ind_and_grad(R, D, qis, qjs, qic, qjc, u, k)
    e, g = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
    # U_ind = U_coul - U_static + U_self: round-off scales with the cancelling sums, not with U_ind
    assert abs(e[0] - e_ref) <= 1e-12 * abs(e_ref) + 1e-14 * (abs(e[1]) + abs(e[2]))
    assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
    t = [O._t(a) for a in (R, qis, qjs, qic, qjc)]
    # sums of +-O(100) kJ/mol pair terms that cancel to a few kJ/mol: compare on the scale of the terms
    scale = 138.935 * float(np.abs(s.q_core).sum() + np.abs(s.q_shell).sum()) ** 2
    assert abs(e[1] - float(O.coulomb_total(t[0], O._t(D), t[1], t[2], t[3], t[4], O._t(u)))) <= 1e-15 * scale
    assert abs(e[2] - float(O.coulomb_static(*t))) <= 1e-15 * scale
    assert e[3] == pytest.approx(float(O.spring_energy(O._t(D), O._t(k))), rel=1e-13)


def test_c_oracle_row_ranges_add_up():
    s = synth.imidazole_box(12, seed=5)
    D = synth.random_displacements(s, seed=6)
    e, g = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
    n = s.nmol * s.natoms
    cut = 5 * s.natoms
    e1, g1 = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=cut)
    e2, g2 = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=cut, row_hi=n)
    assert abs(e1[0] + e2[0] - e[0]) <= 1e-12 * abs(e[0]) + 1e-14 * (abs(e[1]) + abs(e[2]))
    assert np.array_equal(g1.reshape(-1, 3)[:cut], g.reshape(-1, 3)[:cut])
    assert np.array_equal(g2.reshape(-1, 3)[cut:], g.reshape(-1, 3)[cut:])


def test_zero_distance_dropped():
    """polarization.py:33-42: a zero-charge site on top of another molecule's site changes nothing."""
    s = synth.water_box(4, seed=7)
    D = synth.random_displacements(s, seed=8)
    e, _ = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, None)
    pad = np.roll(s.r_core[:, 0, :], -1, axis=0)[:, None, :]
    r2 = np.concatenate([s.r_core, pad], 1)
    z = np.zeros((s.nmol, 1))
    e2, _ = c_oracle.energy_grad(r2, np.concatenate([D, np.zeros((s.nmol, 1, 3))], 1), np.concatenate([s.q_core, z], 1),
                                 np.concatenate([s.q_shell, z], 1), np.concatenate([s.k, z], 1), None)
    assert abs(e2[0] - e[0]) <= 1e-12 * abs(e[0]) + 1e-14 * (abs(e[1]) + abs(e[2]))


def test_long_double_arbiter_agrees_with_both_oracles():
    s = synth.imidazole_box(40, seed=11)
    D = synth.random_displacements(s, seed=12)
    e, _ = c_oracle.energy_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, want_grad=False)
    e_ld = c_oracle.uind_long_double(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u)
    assert e[0] == pytest.approx(e_ld, rel=1e-11)      # fp64 remainder-of-large-sums noise
    w = synth.water_box(27, seed=13)
    Dw = synth.random_displacements(w, seed=14)
    ew, _ = c_oracle.energy_grad(w.r_core, Dw, w.q_core, w.q_shell, w.k, None, want_grad=False)
    assert ew[0] == pytest.approx(c_oracle.uind_long_double(w.r_core, Dw, w.q_core, w.q_shell, w.k, None), rel=1e-11)


@pytest.mark.parametrize("seed", range(8))
def test_c_oracle_vs_dense_random_topologies(seed):
    """Same random-topology generator as the GPU suite: the two CPU oracles agree on every draw."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "_rt", os.path.join(os.path.dirname(__file__), "test_gpu_random_topologies.py"))
    rt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rt)
    r, qc, qs, k, th, D = rt._random_system(1000 + seed)
    Rij, qis, qjs, qic, qjc, u, kk = O.dense_inputs(r, qc, qs, k, th)
    e_ref, g_ref = O.u_ind_and_grad(Rij, D, qis, qjs, qic, qjc, u, kk)
    e, g = c_oracle.energy_grad(r, D, qc, qs, k, th)
    assert abs(e[0] - e_ref) <= 1e-11 * abs(e_ref) + 1e-14 * (abs(e[1]) + abs(e[2]))
    assert np.abs(g - g_ref).max() <= 1e-11 * max(np.abs(g_ref).max(), 1e-9)
