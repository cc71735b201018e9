#!/usr/bin/env python
"""Generate the committed fixtures under tests/golden/ and u_pol_b200/data/monomers.json.

Runs ONLY in the build container (needs /root/reference); the outputs are committed so that
neither the GPU tests, smoke() nor bench.py ever read the reference tree.

  python tests/golden/make_golden.py

What it writes
  u_pol_b200/data/monomers.json  monomer parameters + rigid geometry for the synthetic boxes
  tests/golden/small_systems.npz  compact arrays of the reference's own input systems
  tests/golden/dimers312.npz      the 312 imidazole-dimer geometries + the reference's recorded
                                  minimised U_ind (U_pol/tests/imidazole/induction_data.csv)
  tests/golden/ref_exec.npz       energies / gradients / components computed by the reference's
                                  OWN source (U_pol/polarization.py) executed via oracle/ref_exec.py
"""
import csv
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle import ref_exec, uind_dense            # noqa: E402
from u_pol_b200.topology import build_system       # noqa: E402

REF = ref_exec.REFERENCE_ROOT
BENCH = os.path.join(REF, "benchmarks", "OpenMM")
GOLD = os.path.dirname(os.path.abspath(__file__))
DIMER_DIR = os.path.join(BENCH, "imidazole3", "imidazole", "imidazole", "generated_pdbs")

SMALL = {
    "water": ("water/water.pdb", "water/water.xml"),
    "acnit": ("acnit/acnit.pdb", "acnit/acnit.xml"),
    "imidazole_trimer": ("imidazole/imidazole.pdb", "imidazole/imidazole.xml"),
    "imidazole_dimer10": ("imidazole3/imidazole/imidazole/generated_pdbs/2mer-0+10.pdb", "imidazole3/imidazole3.xml"),
    "imidazole3": ("imidazole3/imidazole3.pdb", "imidazole3/imidazole3.xml"),
}


def monomers():
    w = build_system(os.path.join(BENCH, SMALL["water"][0]), os.path.join(BENCH, SMALL["water"][1]))
    r_oh, ang = 0.09572, np.deg2rad(104.52)
    o = np.zeros(3)
    h1 = r_oh * np.array([np.sin(ang / 2), 0.0, np.cos(ang / 2)])
    h2 = r_oh * np.array([-np.sin(ang / 2), 0.0, np.cos(ang / 2)])
    wts = (0.589781071, 0.2051094645, 0.2051094645)          # water.xml:18
    m = wts[0] * o + wts[1] * h1 + wts[2] * h2
    im = build_system(os.path.join(BENCH, SMALL["imidazole3"][0]), os.path.join(BENCH, SMALL["imidazole3"][1]))
    out = {}
    for name, s, geom in (("water", w, np.array([o, h1, h2, m])), ("imidazole", im, im.r_core[0])):
        out[name] = {
            "site_names": s.site_names,
            "geometry": (geom - geom.mean(0)).tolist(),
            "q_core": s.q_core[0].tolist(), "q_shell": s.q_shell[0].tolist(),
            "alpha": s.alpha[0].tolist(), "thole_u": s.thole_u[0].tolist(),
        }
    os.makedirs(os.path.join(ROOT, "u_pol_b200", "data"), exist_ok=True)
    with open(os.path.join(ROOT, "u_pol_b200", "data", "monomers.json"), "w") as fh:
        json.dump(out, fh, indent=1)


def small_systems():
    out = {}
    for name, (pdb, xml) in SMALL.items():
        for tag, dt in (("", "float64"), ("_f32pos", "float32")):
            s = build_system(os.path.join(BENCH, pdb), os.path.join(BENCH, xml), positions_dtype=dt)
            for f in ("r_core", "q_core", "q_shell", "alpha", "thole_u"):
                out[f"{name}{tag}/{f}"] = getattr(s, f)
    np.savez_compressed(os.path.join(GOLD, "small_systems.npz"), **out)


def dimers():
    rows = {}
    with open(os.path.join(REF, "U_pol", "tests", "imidazole", "induction_data.csv")) as fh:
        for row in csv.DictReader(fh):
            rows[int(float(row["DimerID"]))] = (float(row["SAPT_Induction"]), float(row["MD_Induction"]))
    xml = os.path.join(BENCH, "imidazole3", "imidazole3.xml")
    ids, r, sapt, md = [], [], [], []
    first = None
    for fn in sorted(os.listdir(DIMER_DIR)):
        if not fn.startswith("2mer-0+"):
            continue
        did = int(fn[len("2mer-0+"):-4])
        if did not in rows:
            continue
        s = build_system(os.path.join(DIMER_DIR, fn), xml)
        first = first or s
        ids.append(did); r.append(s.r_core); sapt.append(rows[did][0]); md.append(rows[did][1])
    order = np.argsort(ids)
    np.savez_compressed(
        os.path.join(GOLD, "dimers312.npz"),
        dimer_id=np.array(ids)[order], r_core=np.array(r)[order],
        sapt_induction=np.array(sapt)[order], md_induction=np.array(md)[order],
        q_core=first.q_core, q_shell=first.q_shell, alpha=first.alpha, thole_u=first.thole_u,
    )
    print("dimers:", len(ids))


def ref_exec_vectors():
    """Outputs of the reference's own Uind/Ucoul/Ucoul_static/Uself/make_Sij + reverse-mode grad."""
    import torch
    from u_pol_b200 import synth

    ref = ref_exec.load_reference()
    systems = {}
    for name, (pdb, xml) in SMALL.items():
        systems[name] = build_system(os.path.join(BENCH, pdb), os.path.join(BENCH, xml))
    systems["water_box27"] = synth.water_box(27, seed=11)
    systems["imidazole_box12"] = synth.imidazole_box(12, seed=12)
    out = {}
    for name, s in systems.items():
        R, qis, qjs, qic, qjc, u, k = uind_dense.dense_inputs(s.r_core, s.q_core, s.q_shell, s.k, s.thole_u)
        pol = (s.q_shell != 0)[..., None]
        dsets = {
            "zero": np.zeros_like(s.r_core),
            "pattern": np.where(pol, np.array([0.001, -0.002, 0.0015]), 0.0),
            "random": synth.random_displacements(s, seed=7, sigma=0.002),
            "random_all_sites": np.random.default_rng(8).normal(scale=0.002, size=s.r_core.shape),
        }
        for f in ("r_core", "q_core", "q_shell", "alpha", "thole_u"):
            out[f"{name}/{f}"] = getattr(s, f)
        tt = lambda a: a if np.isscalar(a) else torch.as_tensor(a)
        for dn, D in dsets.items():
            e, g = ref_exec.ref_uind_and_grad(R, D, qis, qjs, qic, qjc, u, k)
            out[f"{name}/{dn}/D"] = D
            out[f"{name}/{dn}/Uind"] = e
            out[f"{name}/{dn}/grad"] = g
            Dt = torch.as_tensor(D)
            out[f"{name}/{dn}/Ucoul"] = float(ref.Ucoul(tt(R), Dt, tt(qis), tt(qjs), tt(qic), tt(qjc), tt(u)))
            out[f"{name}/{dn}/Uself"] = float(ref.Uself(Dt, tt(k)))
        out[f"{name}/Ucoul_static"] = float(ref.Ucoul_static(tt(R), tt(qis), tt(qjs), tt(qic), tt(qjc)))
        if not np.isscalar(u):
            out[f"{name}/make_Sij"] = ref.make_Sij(tt(R), tt(u)).numpy()
        # reference drudeOpt with the SciPy stand-in for jaxopt.BFGS (iterates NOT pinned)
        if s.nmol <= 3:
            d_opt = ref.drudeOpt(tt(R), torch.zeros(s.r_core.size, dtype=torch.float64),
                                 tt(qis), tt(qjs), tt(qic), tt(qjc), tt(u), tt(k))
            out[f"{name}/drudeOpt/d_opt"] = d_opt.numpy()
            out[f"{name}/drudeOpt/Uind"] = float(ref.Uind(tt(R), d_opt, tt(qis), tt(qjs), tt(qic), tt(qjc), tt(u), tt(k)))
            # exact minimum (Newton, autodiff Hessian) from the restated oracle
            dmin, emin, info = uind_dense.newton_minimum(R, np.zeros_like(s.r_core), qis, qjs, qic, qjc, u, k)
            out[f"{name}/Uind_min"] = emin
            out[f"{name}/d_min"] = dmin
            print(name, "drudeOpt U", out[f"{name}/drudeOpt/Uind"], "tight min", emin, info)
    out["ONE_4PI_EPS0"] = ref.ONE_4PI_EPS0
    np.savez_compressed(os.path.join(GOLD, "ref_exec.npz"), **out)


if __name__ == "__main__":
    if not ref_exec.reference_available():
        sys.exit("reference tree not found; fixtures can only be regenerated in the build container")
    monomers()
    small_systems()
    dimers()
    ref_exec_vectors()
    print("done")
