"""Periodic boundaries (Ewald) through the C ABI against oracle/ewald_dense.py.
The reference has no periodic code; the oracle is pinned by tests/test_oracle_ewald.py."""
import numpy as np
import pytest
import torch

from oracle import ewald_dense
from u_pol_b200 import synth
from u_pol_b200._lib import UpolError
from u_pol_b200.engine import UindSystem

pytestmark = pytest.mark.gpu


def _box_for(s, pad=0.35):
    ext = np.ptp(s.r_core.reshape(-1, 3), axis=0)
    return (ext + pad) * np.array([1.0, 1.07, 1.13])


def _compare(s, D, box, kappa=None, tol_e=1e-10, tol_g=1e-10):
    sys_ = UindSystem.from_drude_system(s)
    info = sys_.set_periodic(box, kappa=kappa)
    e, g = sys_.energy_grad(D)
    e_ref, g_ref = ewald_dense.u_ind_pbc_and_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, box, info["kappa"])
    g = g.cpu().numpy().reshape(g_ref.shape)
    # U_ind is the small remainder of row sums of size ~E_STATIC: round-off floor relative to that
    assert abs(float(e[0]) - e_ref) <= tol_e * abs(e_ref) + 1e-14 * abs(float(e[5])), (float(e[0]), e_ref)
    assert np.abs(g - g_ref).max() <= tol_g * np.abs(g_ref).max()
    return sys_, float(e[0]), g


@pytest.mark.parametrize("maker,nmol,seed", [(synth.water_box, 8, 5), (synth.imidazole_box, 6, 7), (synth.water_box, 27, 9)])
def test_energy_and_gradient_match_the_ewald_oracle(maker, nmol, seed):
    s = maker(nmol, seed=seed)
    D = synth.random_displacements(s, seed=seed + 1)
    sys_, e, g = _compare(s, D, _box_for(s))
    assert not g.reshape(-1, 3)[s.q_shell.reshape(-1) == 0].any()
    # gradient only / energy only calls agree with the combined call
    g2 = sys_.energy_grad(D, want_energy=False)[1].cpu().numpy().reshape(g.shape)
    e2 = sys_.energy_grad(D, want_grad=False)[0]
    assert np.abs(g2 - g).max() <= 1e-13 * np.abs(g).max() and abs(float(e2[0]) - e) <= 1e-13 * abs(e)


def test_zero_displacement_and_kappa_independence():
    s = synth.imidazole_box(5, seed=11)
    box = _box_for(s)
    D0 = np.zeros_like(s.r_core)
    sys_, e0, g0 = _compare(s, D0, box)                       # shells on their cores: series branch of erf(x)/x
    assert abs(e0) <= 1e-9                                    # U_ind(d = 0) = 0 up to round-off (kJ/mol)
    D = synth.random_displacements(s, seed=12)
    e_a, g_a = sys_.energy_grad(D)
    sys_.set_periodic(box, kappa=1.3 * sys_.periodic_info()["kappa"])
    e_b, g_b = sys_.energy_grad(D)
    assert abs(float(e_a[0]) - float(e_b[0])) <= 1e-9 * abs(float(e_a[0]))
    assert float((g_a - g_b).abs().max()) <= 1e-9 * float(g_a.abs().max())
    with pytest.raises(UpolError):
        sys_.set_periodic(box, kappa=5.0 / box.min())         # nearest-image real space would be incomplete


def test_lattice_translation_of_whole_molecules_changes_nothing():
    s = synth.water_box(12, seed=3)
    D = synth.random_displacements(s, seed=4)
    box = _box_for(s)
    sys_ = UindSystem.from_drude_system(s)
    sys_.set_periodic(box)
    e, g = sys_.energy_grad(D)
    r2 = s.r_core.copy()
    r2[::2] += box * np.array([1, -1, 2])                     # every other molecule into another image cell
    r2 += np.array([0.3, -5.0, 11.0])                         # and the whole system somewhere else
    sys_.set_positions(r2)
    e2, g2 = sys_.energy_grad(D)
    assert abs(float(e[0]) - float(e2[0])) <= 1e-10 * abs(float(e[0]))
    assert float((g - g2).abs().max()) <= 1e-10 * float(g.abs().max())


def test_large_box_limit_and_switching_back_to_open_boundaries():
    s = synth.water_box(20, seed=21)
    D = synth.random_displacements(s, seed=22)
    sys_ = UindSystem.from_drude_system(s)
    e_open, g_open = sys_.energy_grad(D)
    sys_.set_periodic([150.0, 150.0, 150.0])
    e_big, g_big = sys_.energy_grad(D)
    assert abs(float(e_big[0]) - float(e_open[0])) <= 1e-6 * abs(float(e_open[0]))
    assert float((g_big - g_open).abs().max()) <= 1e-6 * float(g_open.abs().max())
    sys_.set_periodic(_box_for(s))
    e_small = sys_.energy_grad(D)[0]
    assert abs(float(e_small[0]) - float(e_open[0])) > 1e-5 * abs(float(e_open[0]))
    assert sys_.set_periodic(None) is None and sys_.periodic_info() is None
    e_again, g_again = sys_.energy_grad(D)
    assert float(e_again[0]) == float(e_open[0]) and torch.equal(g_again, g_open)


def test_minimiser_in_a_periodic_box():
    s = synth.water_box(27, seed=31)
    box = _box_for(s, pad=0.3)
    sys_ = UindSystem.from_drude_system(s)
    info = sys_.set_periodic(box)
    d, res = sys_.minimize(tol=1e-6)
    assert res["flags"] == 0 and res["gnorm"] <= 1e-6
    e_ref, g_ref = ewald_dense.u_ind_pbc_and_grad(s.r_core, d.cpu().numpy().reshape(s.r_core.shape), s.q_core, s.q_shell,
                                                  s.k, s.thole_u, box, info["kappa"])
    assert abs(res["U_ind"] - e_ref) <= 1e-9 * abs(e_ref)
    assert np.linalg.norm(g_ref) <= 1e-5                      # the oracle agrees that this is a stationary point
    # strictly below the open-boundary-minimised configuration evaluated in the periodic box
    sys_.set_periodic(None)
    d_open, _ = sys_.minimize(tol=1e-6)
    sys_.set_periodic(box)
    e_cross = float(sys_.energy_grad(d_open, want_grad=False)[0][0])
    assert res["U_ind"] < e_cross
    with pytest.raises(UpolError):
        sys_.minimize(precision="mixed")


def test_medium_box_against_the_oracle():
    s = synth.imidazole_box(64, seed=41)                      # 576 sites, 320 Drudes, full 4x4x4 lattice
    D = synth.random_displacements(s, seed=42)
    a = (1.0 / 7.0) ** (1.0 / 3.0)
    box = np.array([4 * a + 0.1] * 3)                         # the liquid's own cell (+ margin for the faces)
    _compare(s, D, box)


def test_broken_or_oversized_molecules_are_refused():
    s = synth.imidazole_box(4, seed=13)
    D = np.zeros_like(s.r_core)
    box = _box_for(s, pad=0.8)
    sys_ = UindSystem.from_drude_system(s)
    sys_.set_periodic(box)
    sys_.energy_grad(D)
    r2 = s.r_core.copy()
    r2[1, 3] += [box[0], 0.0, 0.0]                            # one atom wrapped on its own: molecule no longer whole
    sys_.set_positions(r2)
    with pytest.raises(UpolError):
        sys_.energy_grad(D)
    with pytest.raises(UpolError):
        sys_.minimize()
    sys_.set_positions(s.r_core)
    sys_.energy_grad(D)                                       # recovers with a valid geometry
    ext = np.ptp(s.r_core[0], axis=0).max()
    sys_.set_periodic([1.5 * ext, 10.0, 10.0])                # molecule larger than half the cell
    with pytest.raises(UpolError):
        sys_.energy_grad(D)


@pytest.mark.parametrize("maker,nmol,seed", [(synth.water_box, 8, 51), (synth.imidazole_box, 6, 52), (synth.water_box, 27, 53)])
def test_core_gradient_in_a_periodic_box(maker, nmol, seed):
    """dU_ind/dr_core at fixed d under periodic boundaries (both halves of SURVEY 8f next #3 together)
    against reverse-mode AD of the Ewald oracle."""
    s = maker(nmol, seed=seed)
    D = synth.random_displacements(s, seed=seed + 1)
    box = _box_for(s, pad=0.6)
    sys_ = UindSystem.from_drude_system(s)
    info = sys_.set_periodic(box)
    g = sys_.core_gradient(D).cpu().numpy()
    g_ref = ewald_dense.u_ind_pbc_core_grad(s.r_core, D, s.q_core, s.q_shell, s.k, s.thole_u, box, info["kappa"]).reshape(g.shape)
    assert np.abs(g - g_ref).max() <= 1e-9 * np.abs(g_ref).max()
    assert np.abs(g.sum(0)).max() <= 1e-9 * np.abs(g).max() * len(g)       # no net force on the cell
    # differs from the open-boundary core gradient of the same cluster, and that one is untouched
    sys_.set_periodic(None)
    g_open = sys_.core_gradient(D).cpu().numpy()
    assert np.abs(g_open - g).max() > 1e-6 * np.abs(g).max()
