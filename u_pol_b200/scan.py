"""Batched SAPT dimer-scan driver (SURVEY 8f "next" #1).

Replaces the loop of the reference's ``U_pol/wrapper.py:265-316`` - parse one Psi4 SAPT0 ``.out``,
write a PDB, copy it over ``imidazole3.pdb``, start ``python polarization.py --mol imidazole3`` in a
new interpreter (OpenMM set-up + JAX trace + BFGS per dimer), parse ``log.out`` - by: parse all
``.out`` files, build ONE ``(B, nmol, natoms, 3)`` position tensor, minimise every conformer in a
single launch (``engine.batched_minimize``, one warp per dimer), and write ``induction_data.csv`` in
the reference's format (``wrapper.py:367-373``) so ``plot_sapt_md.py`` works unchanged.

Kept from the reference, because they define the numbers in its recorded CSV:
  * only the FIRST "Dimer HF" geometry block is read, in Bohr (``wrapper.py:85-143``);
  * the induction energy is the kJ/mol column of the "Induction" line of "SAPT Results"
    (``wrapper.py:145-160``);
  * Bohr -> Angstrom with 0.529177210903 and the ``%8.3f`` rounding of the PDB writer
    (``wrapper.py:79,183-185``): the force-field path sees 3-decimal Angstrom coordinates;
  * OpenMM's CPU platform holds positions in float32 (``positions="float32"``, default).
"""
from __future__ import annotations

import os
import re

import numpy as np

BOHR_TO_ANG = 0.529177210903     # wrapper.py:79

_END_MARKERS = ("running in", "nuclear repulsion", "symmetry", "set {", "psi4:")


def parse_sapt_outfile(path):
    """(coords_bohr (n,3) array, induction kJ/mol or None) from a Psi4 SAPT0 output."""
    with open(path) as fh:
        lines = fh.readlines()
    coords, state = [], 0          # 0: looking for "Dimer HF", 1: for the geometry header, 2: reading
    for line in lines:
        if state == 0:
            if "Dimer HF" in line:
                state = 1
        elif state == 1:
            if "Geometry (in Bohr)" in line:
                state = 2
        else:
            low = line.lower()
            if ("monomer" in low and "hf" in low) or "rotational constants" in line or any(m in low for m in _END_MARKERS):
                break
            s = line.strip()
            if not s or s.startswith("Center") or s.startswith("---") or "Gh(" in line:
                continue
            parts = s.split()
            if len(parts) >= 4:
                try:
                    coords.append([float(parts[1]), float(parts[2]), float(parts[3])])
                except ValueError:
                    pass
    induction, in_results = None, False
    for line in lines:
        if "SAPT Results" in line:
            in_results = True
        elif in_results and line.strip().startswith("Induction"):
            parts = line.split()
            if len(parts) >= 6 and parts[-1] == "[kJ/mol]":
                try:
                    induction = float(parts[-2])
                except ValueError:
                    pass
            break
    return np.array(coords, dtype=np.float64).reshape(-1, 3), induction


def pdb_rounded_nm(coords_bohr, positions="float32"):
    """Coordinates as the reference's force-field path sees them: Bohr -> Angstrom, the PDB writer's
    3 decimals (wrapper.py:183-185), Angstrom -> nm, and OpenMM's float32 storage."""
    ang = np.array([[float(f"{v:8.3f}") for v in row] for row in np.asarray(coords_bohr) * BOHR_TO_ANG])
    nm = ang * 0.1
    if positions == "float32":
        nm = nm.astype(np.float32).astype(np.float64)
    return nm


def run_scan(indir, monomer, outfile="induction_data.csv", tol=1e-4, positions="float32", device=None):
    """Parse every ``*+<id>.out`` of ``indir``, minimise all dimers in one launch, write the CSV.

    ``monomer``: dict with ``q_core, q_shell, alpha, thole_u`` for ONE molecule (as in
    ``u_pol_b200/data/monomers.json``); every conformer is nmol = natoms_total / natoms molecules of it.
    Returns (dimer_ids, sapt_induction, md_induction) arrays in file order."""
    from . import _lib as L
    from .engine import batched_minimize
    from .topology import ONE_4PI_EPS0

    natoms = len(monomer["q_core"])
    ids, sapt, geoms = [], [], []
    for fn in sorted(f for f in os.listdir(indir) if f.endswith(".out")):
        m = re.search(r"\+(\d+)\.out$", fn)                    # wrapper.py:268
        if not m:
            continue
        coords, e_ind = parse_sapt_outfile(os.path.join(indir, fn))
        if e_ind is None or coords.size == 0 or coords.shape[0] % natoms != 0:
            continue
        if geoms and coords.shape[0] != geoms[0].shape[0] * natoms:
            continue
        ids.append(int(m.group(1)))
        sapt.append(e_ind)
        geoms.append(pdb_rounded_nm(coords, positions).reshape(-1, natoms, 3))
    if not ids:
        return np.zeros(0, int), np.zeros(0), np.zeros(0)
    r = np.stack(geoms)
    nmol = r.shape[1]
    tile = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), (nmol,) + np.shape(v))
    qc, qs, al = tile(monomer["q_core"]), tile(monomer["q_shell"]), tile(monomer["alpha"])
    k = np.where(al == 0, 0.0, ONE_4PI_EPS0 * qs**2 / np.where(al == 0, 1.0, al))
    _, e, _ = batched_minimize(r, qc, qs, k, tile(monomer["thole_u"]), tol=tol, device=device)
    e = e.cpu().numpy()
    md = e[:, 0].copy()
    ids, sapt = np.array(ids), np.array(sapt)
    # per-system UPOL_FLAG_* bits (column 6): a minimisation that went non-finite or did not converge must not
    # become a plausible-looking CSV row - it is written as nan and counted
    flags = e[:, 6].astype(np.int64)
    bad = (flags & (L.FLAG_NONFINITE | L.FLAG_NOT_CONVERGED)) != 0
    if bad.any():
        import warnings

        md[bad] = np.nan
        warnings.warn(f"run_scan: {int(bad.sum())} of {len(md)} dimers did not minimise cleanly "
                      f"(ids {ids[bad][:10].tolist()}{'...' if bad.sum() > 10 else ''}); written as nan", RuntimeWarning)
    if outfile:
        np.savetxt(outfile, np.column_stack([sapt, md, ids]), delimiter=",",
                   header="SAPT_Induction,MD_Induction,DimerID", comments="")     # wrapper.py:367-373
    return ids, sapt, md
