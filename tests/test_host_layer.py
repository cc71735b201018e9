"""CPU: host-side logic - PDB/XML reader, reference-layout array builders, synthetic generators,
dense->compact reduction used by the drop-in API."""
import os

import numpy as np
import pytest

from oracle import ref_exec as rx
from u_pol_b200 import synth, util
from u_pol_b200.engine import shard_bounds, thole_pairs_from_blocks
from u_pol_b200.topology import ONE_4PI_EPS0, DrudeSystem, build_system

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BENCH = os.path.join(rx.REFERENCE_ROOT, "benchmarks", "OpenMM")
needs_ref = pytest.mark.skipif(not rx.reference_available(), reason="reference input files not present")


def test_constant():
    assert ONE_4PI_EPS0 == 138.93545764438198


@needs_ref
@pytest.mark.parametrize("name,pdb,xml", [
    ("water", "water/water.pdb", "water/water.xml"),
    ("acnit", "acnit/acnit.pdb", "acnit/acnit.xml"),
    ("imidazole_trimer", "imidazole/imidazole.pdb", "imidazole/imidazole.xml"),
])
def test_reader_reproduces_committed_fixture(name, pdb, xml):
    z = np.load(os.path.join(GOLD, "small_systems.npz"))
    s = build_system(os.path.join(BENCH, pdb), os.path.join(BENCH, xml))
    for f in ("r_core", "q_core", "q_shell", "alpha", "thole_u"):
        assert np.array_equal(getattr(s, f), z[f"{name}/{f}"])


@needs_ref
def test_reader_semantics():
    w = build_system(os.path.join(BENCH, "water/water.pdb"), os.path.join(BENCH, "water/water.xml"))
    assert w.site_names == ["O", "H1", "H2", "M"] and w.r_core.shape == (2, 4, 3)
    assert not w.has_screened_pairs                      # one Drude per molecule (util.py:275-277)
    assert w.q_shell[0].tolist() == [-1.71636, 0.0, 0.0, 0.0]
    # M site: average3 of O, H1, H2 (water.xml:18); OpenMM printed -0.017725981771945953 (raw_inputs.txt:33)
    assert w.r_core[0, 3, 0] == pytest.approx(-0.01772598, abs=1e-7)
    a = build_system(os.path.join(BENCH, "acnit/acnit.pdb"), os.path.join(BENCH, "acnit/acnit.xml"))
    assert int((a.thole_u[0] != 0).sum()) == 6           # 3 screened pairs per molecule, symmetric
    i, j = a.site_names.index("CT"), a.site_names.index("CZ")
    assert a.thole_u[0, i, j] == pytest.approx(18.19093587791603, rel=1e-14)
    assert a.k[0, i] == pytest.approx(93709.93356256426, rel=1e-14)
    im = build_system(os.path.join(BENCH, "imidazole3/imidazole3.pdb"), os.path.join(BENCH, "imidazole3/imidazole3.xml"))
    assert int((im.thole_u[0] != 0).sum()) == 20 and im.natoms == 9 and int((im.q_shell[0] != 0).sum()) == 5
    f32 = build_system(os.path.join(BENCH, "acnit/acnit.pdb"), os.path.join(BENCH, "acnit/acnit.xml"), positions_dtype="float32")
    assert np.array_equal(f32.r_core, a.r_core.astype(np.float32).astype(np.float64))


def test_util_dense_layout_matches_reference_shapes():
    s = synth.imidazole_box(3, seed=1)
    Rij, Dij = util.get_Rij_Dij(s)
    assert Rij.shape == (3, 3, 9, 9, 3) and Dij.shape == (3, 9, 3) and not Dij.any()
    assert np.array_equal(Rij[1, 2, 3, 4], s.r_core[2, 4] - s.r_core[1, 3])          # util.py:78-81
    Qi_core, Qi_shell, Qj_core, Qj_shell = util.get_QiQj(s)
    assert Qi_core.shape == (3, 1, 9, 1) and Qj_shell.shape == (1, 3, 1, 9)
    k, u = util.get_pol_params(s)
    assert k.shape == (3, 9) and u.shape == (3, 3, 9, 9)
    assert not u[0, 1].any() and np.array_equal(u[1, 1], s.thole_u[1])
    w = synth.water_box(2, seed=1)
    assert util.get_pol_params(w)[1] == 0.0                                             # util.py:277
    d = util.get_Dij(np.ones((1, 2, 3)), np.array([[[0.5, 0.5, 0.5], [0.0, 0.0, 0.0]]]))
    assert d[0, 0].tolist() == [0.5, 0.5, 0.5] and d[0, 1].tolist() == [0.0, 0.0, 0.0]  # util.py:37-42


def test_dense_to_compact_reduction_roundtrip():
    from u_pol_b200.polarization import _compact
    s = synth.imidazole_box(4, seed=2)
    Rij, _ = util.get_Rij_Dij(s)
    Qi_core, Qi_shell, _, _ = util.get_QiQj(s)
    k, u = util.get_pol_params(s)
    r, qc, qs, kk, th = _compact(Rij, Qi_shell, Qi_core, u, k)
    assert np.allclose(r, s.r_core - s.r_core[0, 0], atol=1e-15)      # translation-invariant
    assert np.array_equal(qc, s.q_core) and np.array_equal(qs, s.q_shell) and np.array_equal(th, s.thole_u)
    a, b, uu = thole_pairs_from_blocks(th, s.natoms)
    assert len(a) == 4 * 10 and np.all(a < b) and np.all(a // 9 == b // 9) and np.all(uu > 0)
    with pytest.raises(ValueError):
        _compact(Rij, Qi_shell, Qi_core, 1.0, k)
    with pytest.raises(ValueError):
        _compact(Rij[0], Qi_shell, Qi_core, u, k)


@pytest.mark.parametrize("maker,nmol,seed,dmin", [(synth.water_box, 64, 3, 0.18), (synth.imidazole_box, 40, 3, 0.20),
                                                  (synth.water_box, 3000, 2, 0.18), (synth.water_box, 887, 105386, 0.18),
                                                  (synth.imidazole_box, 1500, 11, 0.20)])
def test_synthetic_boxes_respect_min_distance(maker, nmol, seed, dmin):
    # the larger cases include seeds whose stubborn molecules once escaped the neighbour test
    from scipy.spatial import cKDTree
    s = maker(nmol, seed=seed)
    r = s.r_core.reshape(-1, 3)
    mol = np.repeat(np.arange(nmol), s.natoms)
    close = cKDTree(r).query_pairs(dmin - 1e-12, output_type="ndarray")
    assert not np.any(mol[close[:, 0]] != mol[close[:, 1]])
    seed = 3 if nmol > 100 else seed
    nmol = min(nmol, 64)
    s = maker(nmol, seed=3)
    assert np.array_equal(maker(nmol, seed=3).r_core, s.r_core)        # deterministic
    dd = synth.random_displacements(s, seed=1)
    assert not dd[s.q_shell == 0].any() and dd[s.q_shell != 0].any()


def test_perturbed_dimers():
    z = np.load(os.path.join(GOLD, "dimers312.npz"))
    out = synth.perturbed_dimers(z["r_core"], 400, seed=5)
    assert out.shape == (400, 2, 9, 3) and np.array_equal(out[:312], z["r_core"])
    dmin = np.linalg.norm(out[:, 0, :, None] - out[:, 1, None], axis=-1).reshape(400, -1).min(1)
    assert dmin[312:].min() >= 0.25


def test_shard_bounds_partition():
    for nmol, ranks in ((20000, 8), (7, 3), (5, 8)):
        b = [shard_bounds(nmol, ranks, r) for r in range(ranks)]
        assert b[0][0] == 0 and b[-1][1] == nmol and all(b[i][1] == b[i + 1][0] for i in range(ranks - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
