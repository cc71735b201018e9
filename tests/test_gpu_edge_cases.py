"""GPU edge cases: empty / ragged / degenerate inputs (SURVEY section 7 hard parts, 8b semantics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import uind_dense  # noqa: E402  (checker only)
from u_pol_b200 import synth  # noqa: E402
from u_pol_b200.engine import UindSystem  # noqa: E402
from u_pol_b200.topology import ONE_4PI_EPS0, DrudeSystem  # noqa: E402


def test_no_polarisable_site():
    """q_shell = 0 everywhere: U_ind = 0, zero gradient, the minimiser returns at once."""
    s = synth.water_box(8, seed=1)
    sys_ = UindSystem.from_compact(s.r_core, s.q_core, np.zeros_like(s.q_shell), np.zeros_like(s.q_shell), None)
    assert sys_.ndrude == 0
    e, g = sys_.energy_grad(np.random.default_rng(0).normal(size=s.r_core.shape))
    assert float(e[0]) == 0.0 and not g.any()
    d, info = sys_.minimize()
    assert info["iterations"] == 0 and info["flags"] == 0 and not d.any()


def test_single_molecule_has_only_intra_and_self_terms():
    s = synth.imidazole_box(1, seed=2)
    D = synth.random_displacements(s, seed=3)
    R, qis, qjs, qic, qjc, u, k = uind_dense.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
    e_ref, g_ref = uind_dense.u_ind_and_grad(R, D, qis, qjs, qic, qjc, u, k)
    e, g = UindSystem.from_drude_system(s).energy_grad(D)
    assert float(e[1]) == 0.0                                     # no inter-molecular pair
    assert float(e[0]) == pytest.approx(e_ref, rel=1e-12)
    assert np.abs(g.cpu().numpy().reshape(g_ref.shape) - g_ref).max() <= 1e-12 * np.abs(g_ref).max()


def test_ragged_molecules_flat_api_vs_padded_oracle():
    """Molecules of different sizes (water: 4 sites, imidazole: 9) through the flat API.  The reference
    layout is rectangular (util.py:74); the oracle gets zero-charge padding sites, which contribute
    nothing (polarization.py:33-42) - SURVEY section 7, 'rectangular-molecule assumption'."""
    w, im = synth.load_monomer("water"), synth.load_monomer("imidazole")
    rng = np.random.default_rng(5)
    mols = []
    for m in range(6):
        mono = w if m % 2 == 0 else im
        shift = np.array([0.9 * m, 0.3 * (m % 3), 0.2 * (m % 2)])
        mols.append((mono, mono["geometry"] + shift))
    # flat arrays
    r = np.concatenate([g for _, g in mols])
    qc = np.concatenate([m["q_core"] for m, _ in mols])
    qs = np.concatenate([m["q_shell"] for m, _ in mols])
    al = np.concatenate([m["alpha"] for m, _ in mols])
    k = np.where(al == 0, 0.0, ONE_4PI_EPS0 * qs**2 / np.where(al == 0, 1.0, al))
    mol_id = np.concatenate([np.full(len(g), i, dtype=np.int32) for i, (_, g) in enumerate(mols)])
    pa, pb, pu, off = [], [], [], 0
    for mono, g in mols:
        a, b = np.nonzero(np.triu(mono["thole_u"], 1))
        pa += list(off + a); pb += list(off + b); pu += list(mono["thole_u"][a, b])
        off += len(g)
    d = np.where((qs != 0)[:, None], rng.normal(scale=0.002, size=r.shape), 0.0)
    sys_ = UindSystem(r, qc, qs, k, mol_id, np.array(pa, np.int32), np.array(pb, np.int32), np.array(pu))
    e, g = sys_.energy_grad(d)
    # padded rectangular copy for the oracle (9 sites per molecule)
    nm, na = len(mols), 9
    R = np.zeros((nm, na, 3)); QC = np.zeros((nm, na)); QS = np.zeros((nm, na)); KK = np.zeros((nm, na))
    TH = np.zeros((nm, na, na)); DD = np.zeros((nm, na, 3))
    off = 0
    for i, (mono, gm) in enumerate(mols):
        n = len(gm)
        R[i, :n] = gm; R[i, n:] = gm[0]                      # padding sits on the molecule's first atom
        QC[i, :n] = mono["q_core"]; QS[i, :n] = mono["q_shell"]; KK[i, :n] = k[off:off + n]
        TH[i, :n, :n] = mono["thole_u"]; DD[i, :n] = d[off:off + n]
        off += n
    Rij, qis, qjs, qic, qjc, u, kk = uind_dense.dense_inputs(R, QC, QS, KK, TH)
    e_ref, g_ref = uind_dense.u_ind_and_grad(Rij, DD, qis, qjs, qic, qjc, u, kk)
    assert float(e[0]) == pytest.approx(e_ref, rel=1e-10)
    g = g.cpu().numpy()
    off = 0
    for i, (_, gm) in enumerate(mols):
        n = len(gm)
        assert np.abs(g[off:off + n] - g_ref[i, :n]).max() <= 1e-10 * np.abs(g_ref).max()
        off += n
    _, info = sys_.minimize(tol=1e-6)
    assert info["flags"] == 0 and info["U_ind"] < 0


def test_unsorted_molecule_ids_are_rejected():
    from u_pol_b200._lib import UpolError
    s = synth.water_box(3, seed=1)
    r, qc, qs, k, mol = s.flat()
    with pytest.raises(UpolError):
        UindSystem(r, qc, qs, k, mol[::-1].copy())


def test_polarisation_catastrophe_is_flagged_not_hung():
    """Two unscreened Drude sites of different molecules 0.05 nm apart: U_ind is unbounded below
    (SURVEY section 7).  The minimiser must come back with a flag, not hang or crash."""
    s = synth.water_box(2, seed=4)
    r = s.r_core.copy()
    r[1] += (r[0, 0] - r[1, 0]) + np.array([0.05, 0.0, 0.0])
    sys_ = UindSystem.from_compact(r, s.q_core, s.q_shell, s.k, None)
    d, info = sys_.minimize(tol=1e-4, maxiter=50)
    assert info["flags"] != 0 or info["U_ind"] < -1e3
    assert info["evaluations"] <= 50 * 14 + 20


def test_displacements_on_non_polarisable_sites_are_ignored():
    s = synth.imidazole_box(5, seed=6)
    D = synth.random_displacements(s, seed=7)
    D2 = D + np.where((s.q_shell == 0)[..., None], 0.01, 0.0)
    sys_ = UindSystem.from_drude_system(s)
    e1, g1 = sys_.energy_grad(D)
    e2, g2 = sys_.energy_grad(D2)
    assert float(e1[0]) == float(e2[0]) and torch.equal(g1, g2)


@pytest.mark.parametrize("maker,nmol,seed", [(synth.water_box, 20, 71), (synth.imidazole_box, 14, 72)])
def test_core_gradient_vs_autodiff_of_the_oracle(maker, nmol, seed):
    """dU_ind/dr_core at fixed d (new) against reverse-mode AD of the dense oracle
    with Rij rebuilt from r inside the graph."""
    s = maker(nmol, seed=seed)
    D = synth.random_displacements(s, seed=seed + 1)
    r = torch.tensor(s.r_core, dtype=torch.float64, requires_grad=True)
    Rij = r[None, :, None, :, :] - r[:, None, :, None, :]
    _, qis, qjs, qic, qjc, u, k = uind_dense.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
    e = uind_dense.u_ind(Rij, D, qis, qjs, qic, qjc, u, k)
    (g_ref,) = torch.autograd.grad(e, r)
    g_ref = g_ref.numpy()
    g = UindSystem.from_drude_system(s).core_gradient(D).cpu().numpy().reshape(g_ref.shape)
    assert np.abs(g - g_ref).max() <= 1e-9 * np.abs(g_ref).max()
    # translation invariance: the core gradient plus the shells' share sums to zero net force
    assert np.abs(g.reshape(-1, 3).sum(0)).max() <= 1e-9 * np.abs(g).max() * len(g.reshape(-1, 3))


def test_host_buffer_minimise_and_device_positions():
    """upol_minimize_host (numpy in/out through the C ABI) and upol_system_set_positions_dev."""
    import ctypes as C
    from u_pol_b200 import _lib as L

    s = synth.imidazole_box(10, seed=81)
    sys_ = UindSystem.from_drude_system(s)
    d_dev, info = sys_.minimize(tol=1e-6)
    d = np.zeros(3 * sys_.nsite)
    o = L.MinOptions()
    L.lib().upol_min_options_default(C.byref(o))
    o.tol = 1e-6
    res = L.MinResult()
    rc = L.lib().upol_minimize_host(sys_._h, d.ctypes.data, C.byref(o), C.byref(res))
    assert rc == 0 and res.flags == 0 and res.gnorm <= 1e-6
    assert res.energy[0] == pytest.approx(info["U_ind"], rel=1e-12)
    assert np.abs(d.reshape(-1, 3) - d_dev.cpu().numpy()).max() <= 1e-12
    # move the whole box by 50 nm through the device-pointer entry: box centre / fixed-point grid follow
    r2 = torch.as_tensor(s.r_core.reshape(-1, 3) + 50.0).cuda()
    sys_.set_positions(r2)
    e2, g2 = sys_.energy_grad(d_dev)
    assert float(e2[0]) == pytest.approx(info["U_ind"], rel=1e-9)
    gm = sys_.energy_grad(d_dev, precision="mixed", want_energy=False)[1]
    D = synth.random_displacements(s, seed=82)
    g64 = sys_.energy_grad(D)[1]
    gmx = sys_.energy_grad(D, precision="mixed", want_energy=False)[1]
    assert float((gmx - g64).abs().max()) <= 1e-5 * float(g64.abs().max())


def test_plain_c_program_through_the_c_abi(tmp_path):
    """examples/c_api_example.c: a C99 program (no Python, no torch) creates a system, evaluates and
    minimises through the host-buffer entry points."""
    import os
    import shutil
    import subprocess
    from u_pol_b200 import _lib as L

    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "c_api_example")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "examples", "c_api_example.c"), "-L", libdir, "-lupol_b200",
                    f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    line = [l for l in run.stdout.splitlines() if l.startswith("minimised U_ind")][0]
    u = float(line.split("=")[1].split()[0])
    # same system through the Python mirror
    K = 138.93545764438198
    r = np.array([[0.0, 0.0, 0.0], [0.0757, 0.0, 0.0586], [-0.0757, 0.0, 0.0586], [0.0, 0.0, 0.0240],
                  [0.30, 0.0, 0.0], [0.3757, 0.0, 0.0586], [0.2243, 0.0, 0.0586], [0.30, 0.0, 0.0240]]).reshape(2, 4, 3)
    qc = np.tile([1.71636, 0.55733, 0.55733, -1.11466], (2, 1))
    qs = np.tile([-1.71636, 0.0, 0.0, 0.0], (2, 1))
    k = np.where(qs != 0, K * qs**2 / 0.000978253, 0.0)
    _, info = UindSystem.from_compact(r, qc, qs, k, None).minimize(tol=1e-4)
    assert u == pytest.approx(info["U_ind"], rel=1e-8) and u < 0


def test_one_pair_kernel_launch_per_evaluation():
    """An evaluation is prepare -> ONE pair kernel -> finalize (its last block sums the energy vector, and it writes
    the (nsite,3) gradient itself): three launches.  The fp64 pass forms the static potential in the same launch (no
    separate static pass), and a mixed call launches the difference-form kernel only (no fp64 pair kernel, no static
    pass).  A row shard keeps the separate zero-filling scatter (four launches)."""
    s = synth.imidazole_box(20, seed=91)
    D = synth.random_displacements(s, seed=92)
    sys_ = UindSystem.from_drude_system(s)
    d_dev = torch.as_tensor(D.reshape(-1)).cuda()
    n0 = sys_.launch_count()
    e64, g64 = sys_.energy_grad(d_dev)                       # static not cached yet: fused into the pass
    n1 = sys_.launch_count()
    assert n1 - n0 == 3
    em, gm = sys_.energy_grad(d_dev, precision="mixed")
    assert sys_.launch_count() - n1 == 3
    assert abs(float(em[0]) - float(e64[0])) <= 1e-5 * abs(float(e64[0]))
    assert float((gm - g64).abs().max()) <= 1e-5 * float(g64.abs().max())
    # the static potential was kept by the fused pass: the second fp64 evaluation gives the same bits
    e2, g2 = sys_.energy_grad(d_dev)
    assert torch.equal(g2, g64) and abs(float(e2[0]) - float(e64[0])) <= 1e-13 * abs(float(e64[0]))
    assert float(e2[5]) == float(e64[5]) != 0.0              # E_STATIC reported either way
    # non-polarisable sites: exact zeros written by the finalize kernel; a shard zero-fills the other rows' sites too
    nonpol = torch.as_tensor((s.q_shell.reshape(-1) == 0)).cuda()
    assert bool((g64[nonpol] == 0).all()) and bool((g64[~nonpol].abs().sum(dim=1) > 0).all())
    sys_.set_row_range(3, 11)
    n2 = sys_.launch_count()
    _, gs = sys_.energy_grad(d_dev)
    assert sys_.launch_count() - n2 == 4
    na = s.natoms
    assert torch.equal(gs[3 * na:11 * na], g64[3 * na:11 * na]) and bool((gs[:3 * na] == 0).all()) and bool((gs[11 * na:] == 0).all())


def test_dense_wrappers_reuse_the_device_system():
    """drudeOpt followed by Uind on the same dense inputs (the reference's main(), polarization.py:246-257)
    builds ONE device system."""
    from u_pol_b200 import polarization as pol
    from u_pol_b200 import util

    s = synth.imidazole_box(3, seed=93)
    Rij, _ = util.get_Rij_Dij(s)
    Qi_core, Qi_shell, Qj_core, Qj_shell = util.get_QiQj(s)
    k, u_scale = util.get_pol_params(s)
    pol.clear_cache()
    d_opt = pol.drudeOpt(Rij, np.zeros(3 * s.nmol * s.natoms), Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k)
    assert len(pol._CACHE) == 1
    first = next(iter(pol._CACHE.values()))
    e = pol.Uind(Rij, d_opt, Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k)
    assert len(pol._CACHE) == 1 and next(iter(pol._CACHE.values())) is first
    assert float(e) < 0
    us = float(pol.Uself(d_opt, k))
    assert us == pytest.approx(0.5 * float((np.asarray(k).reshape(-1, 1) * d_opt.cpu().numpy().reshape(-1, 3) ** 2).sum()), rel=1e-12)
    pol.clear_cache()
    assert not pol._CACHE


def test_sharded_host_entry_single_rank_pageable_and_pinned_buffers():
    """upol_uind_energy_grad_host_sharded with one rank = set_positions + energy_grad_host in one call; pageable
    numpy buffers are staged through the library's own page-locked memory, page-locked ones are used directly."""
    import torch

    s = synth.imidazole_box(40, seed=21)
    D = synth.random_displacements(s, seed=22)
    sysm = UindSystem.from_drude_system(s)
    e_ref, g_ref = sysm.energy_grad(D)
    r2 = s.r_core + 0.013                                  # a translated copy: same energy, new geometry upload
    for pinned in (False, True):
        mk = (lambda a: torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).pin_memory().numpy()) if pinned else \
             (lambda a: np.ascontiguousarray(a.reshape(-1)))
        e, g, (lo, hi) = sysm.energy_grad_host_sharded(mk(np.asarray(D)), r_core=mk(r2))
        assert (lo, hi) == (0, s.nmol * s.natoms)
        assert abs(e[0] - float(e_ref[0])) <= 1e-10 * abs(float(e_ref[0]))
        assert np.abs(g - g_ref.cpu().numpy()).max() <= 1e-10 * float(g_ref.abs().max())
        em, gm, _ = sysm.energy_grad_host_sharded(mk(np.asarray(D)), precision="mixed")     # geometry kept
        assert abs(em[0] - float(e_ref[0])) <= 1e-5 * abs(float(e_ref[0]))
