"""OpenMM-free host input layer: PDB + Drude force-field XML -> compact site arrays.

Host-side replacement for what the reference obtains from OpenMM objects in
``U_pol/util.py:45-279`` (``get_Rij_Dij``, ``get_QiQj``, ``get_pol_params``) after
``setup_openmm`` (``util.py:282-340``) has built the System.  Topology parsing and Drude
assignment stay on the host (north star); this module produces the *compact* arrays the
CUDA path consumes and, through ``u_pol_b200.util``, the dense broadcast tensors of the
reference layout.

Rules reproduced (reference file:line, or OpenMM behaviour the reference relies on):

* sites of a residue = its PDB atoms in PDB order, then the template's extra non-Drude
  particles (virtual sites) in template order; Drude particles are dropped
  (``util.py:59-72``).
* core charge = ``NonbondedForce`` charge of the atom (``util.py:121-122,134``); shell
  charge = ``DrudeForce`` particle charge of the Drude attached to it, else 0
  (``util.py:124-132``); polarisability likewise (``util.py:193-250``).
* ``k = ONE_4PI_EPS0 * q_shell**2 / alpha``, 0 where ``alpha == 0`` (``util.py:264-268``).
* screened (Thole) pairs: OpenMM's Drude generator registers one screened pair for every
  Drude-Drude nonbonded *exclusion*, i.e. Drude pairs whose parents are 1-2 or 1-3 bonded,
  or 1-4 bonded when the scaled 1-4 charge product and epsilon are both zero; the pair's
  thole is ``thole_1 + thole_2``; ``u = thole / (alpha_1 alpha_2)**(1/6)``
  (``util.py:206-245``).
* ``average3`` virtual sites are placed as ``w1 r1 + w2 r2 + w3 r3``.
* positions are Angstrom in the PDB, nm in the arrays.  OpenMM's CPU platform (the one the
  reference uses, ``util.py:289``) holds positions in float32; ``positions_dtype="float32"``
  reproduces that rounding (see ``benchmarks/OpenMM/water/raw_inputs.txt:33``).
"""
from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from collections import deque

import numpy as np

M_PI = 3.14159265358979323846
E_CHARGE = 1.602176634e-19
AVOGADRO = 6.02214076e23
EPSILON0 = 1e-6 * 8.8541878128e-12 / (E_CHARGE * E_CHARGE * AVOGADRO)
#: Coulomb constant in kJ mol^-1 nm e^-2 (``polarization.py:19-25``); 138.93545764438198
ONE_4PI_EPS0 = 1 / (4 * M_PI * EPSILON0)


@dataclasses.dataclass
class ResidueTemplate:
    name: str
    atom_names: list          # every particle of the template, template order
    atom_types: list
    bonds: list               # pairs of template indices
    virtual_sites: dict       # template index -> (kind, [template indices], [weights])


@dataclasses.dataclass
class ForceFieldParams:
    templates: list
    charge: dict              # atom type -> NonbondedForce charge
    epsilon: dict             # atom type -> LJ epsilon
    coulomb14scale: float
    lj14scale: float
    drude: dict               # parent type -> (drude type, charge, polarizability, thole)
    drude_types: set


@dataclasses.dataclass
class DrudeSystem:
    """Compact description of a homogeneous Drude system (one residue type, as the
    reference assumes: ``util.py:74,209-212``)."""

    r_core: np.ndarray        # (nmol, natoms, 3) nm
    q_core: np.ndarray        # (nmol, natoms)
    q_shell: np.ndarray       # (nmol, natoms)  0 on non-polarisable sites
    alpha: np.ndarray         # (nmol, natoms)  nm^3, 0 on non-polarisable sites
    thole_u: np.ndarray       # (nmol, natoms, natoms) u_ab, 0 where not a screened pair
    site_names: list

    @property
    def nmol(self):
        return self.r_core.shape[0]

    @property
    def natoms(self):
        return self.r_core.shape[1]

    @property
    def k(self):
        """Spring constants incl. the Coulomb constant (``util.py:267-268``)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            k = ONE_4PI_EPS0 * self.q_shell**2 / self.alpha
        return np.where(self.alpha == 0.0, 0.0, k)

    @property
    def has_screened_pairs(self):
        return bool(np.any(self.thole_u != 0.0))

    def flat(self):
        """Flat site arrays for the scalable API: r (N,3), qc, qs, k (N,), mol_id (N,) int32."""
        n = self.nmol * self.natoms
        mol = np.repeat(np.arange(self.nmol, dtype=np.int32), self.natoms)
        return (
            np.ascontiguousarray(self.r_core.reshape(n, 3)),
            np.ascontiguousarray(self.q_core.reshape(n)),
            np.ascontiguousarray(self.q_shell.reshape(n)),
            np.ascontiguousarray(self.k.reshape(n)),
            mol,
        )


def read_forcefield(ff_file) -> ForceFieldParams:
    root = ET.parse(ff_file).getroot()
    templates = []
    for res in root.iter("Residue"):
        names, types = [], []
        for a in res.findall("Atom"):
            names.append(a.get("name"))
            types.append(a.get("type"))

        def _idx(v):
            return int(v) if v.lstrip("-").isdigit() else names.index(v)

        bonds = []
        for b in res.findall("Bond"):
            if b.get("from") is not None:
                bonds.append((_idx(b.get("from")), _idx(b.get("to"))))
            else:
                bonds.append((names.index(b.get("atomName1")), names.index(b.get("atomName2"))))
        vsites = {}
        for v in res.findall("VirtualSite"):
            kind = v.get("type")
            site = _idx(v.get("index")) if v.get("index") is not None else names.index(v.get("siteName"))
            if kind == "average2":
                parents, weights = ["atom1", "atom2"], ["weight1", "weight2"]
            elif kind == "average3":
                parents, weights = ["atom1", "atom2", "atom3"], ["weight1", "weight2", "weight3"]
            else:
                raise NotImplementedError(f"virtual site type {kind!r}")
            vsites[site] = (kind, [_idx(v.get(p)) for p in parents], [float(v.get(w)) for w in weights])
        templates.append(ResidueTemplate(res.get("name"), names, types, bonds, vsites))

    nb = root.find("NonbondedForce")
    charge, epsilon = {}, {}
    for a in nb.findall("Atom"):
        charge[a.get("type")] = float(a.get("charge"))
        epsilon[a.get("type")] = float(a.get("epsilon"))
    drude, drude_types = {}, set()
    df = root.find("DrudeForce")
    if df is not None:
        for p in df.findall("Particle"):
            drude[p.get("type2")] = (
                p.get("type1"),
                float(p.get("charge")),
                float(p.get("polarizability")),
                float(p.get("thole")),
            )
            drude_types.add(p.get("type1"))
    return ForceFieldParams(
        templates, charge, epsilon,
        float(nb.get("coulomb14scale")), float(nb.get("lj14scale")),
        drude, drude_types,
    )


def read_pdb(pdb_file):
    """Return a list of residues, each a list of (atom_name, xyz_angstrom)."""
    residues, key = [], None
    with open(pdb_file) as fh:
        for line in fh:
            if not line.startswith(("ATOM", "HETATM")):
                continue
            name = line[12:16].strip()
            reskey = (line[17:22], line[22:26])
            xyz = (float(line[30:38]), float(line[38:46]), float(line[46:54]))
            if reskey != key:
                residues.append([])
                key = reskey
            residues[-1].append((name, xyz))
    return residues


def _bond_distances(n, bonds):
    dist = np.full((n, n), 99, dtype=np.int64)
    adj = [[] for _ in range(n)]
    for a, b in bonds:
        adj[a].append(b)
        adj[b].append(a)
    for s in range(n):
        dist[s, s] = 0
        dq = deque([s])
        while dq:
            u = dq.popleft()
            for v in adj[u]:
                if dist[s, v] == 99:
                    dist[s, v] = dist[s, u] + 1
                    dq.append(v)
    return dist


def build_system(pdb_file, ff_file, positions_dtype="float64") -> DrudeSystem:
    """PDB + force-field XML -> :class:`DrudeSystem` (no OpenMM)."""
    ff = read_forcefield(ff_file)
    residues = read_pdb(pdb_file)
    if not residues:
        raise ValueError(f"no atoms in {pdb_file}")

    def _template_for(res_atoms):
        names = [a for a, _ in res_atoms]
        for t in ff.templates:
            real = [
                n for i, n in enumerate(t.atom_names)
                if t.atom_types[i] not in ff.drude_types and i not in t.virtual_sites
            ]
            if sorted(real) == sorted(names):
                return t
        raise ValueError(f"no residue template matches atoms {names}")

    r_all, qc_all, qs_all, al_all, u_all, site_names = [], [], [], [], [], None
    for res_atoms in residues:
        t = _template_for(res_atoms)
        tnames = t.atom_names
        pos_by_tidx = {}
        order = []                       # template indices of the sites, reference order
        for name, xyz in res_atoms:      # PDB order first (util.py:61-70)
            ti = tnames.index(name)
            p = np.asarray(xyz, dtype=np.float64) * 0.1
            if positions_dtype == "float32":
                p = p.astype(np.float32)
            pos_by_tidx[ti] = p
            order.append(ti)
        for ti in sorted(t.virtual_sites):   # extra particles appended in template order
            _, parents, weights = t.virtual_sites[ti]
            acc = sum(w * pos_by_tidx[p] for w, p in zip(weights, parents))
            if positions_dtype == "float32":
                acc = np.asarray(
                    sum(np.float32(w) * pos_by_tidx[p] for w, p in zip(weights, parents)),
                    dtype=np.float32,
                )
            pos_by_tidx[ti] = acc
            order.append(ti)
        nsite = len(order)
        r = np.array([np.asarray(pos_by_tidx[ti], dtype=np.float64) for ti in order])
        qc = np.array([ff.charge[t.atom_types[ti]] for ti in order])
        qs = np.zeros(nsite)
        al = np.zeros(nsite)
        th = np.zeros(nsite)
        for s, ti in enumerate(order):
            d = ff.drude.get(t.atom_types[ti])
            if d is not None and d[0] in t.atom_types:
                qs[s], al[s], th[s] = d[1], d[2], d[3]
        # screened pairs from the bond graph of the parents
        dist = _bond_distances(len(tnames), t.bonds)
        u = np.zeros((nsite, nsite))
        for a in range(nsite):
            for b in range(a + 1, nsite):
                if al[a] == 0.0 or al[b] == 0.0:
                    continue
                nb = dist[order[a], order[b]]
                excluded = nb <= 2
                if nb == 3:
                    # a 1-4 exception counts as an exclusion when its scaled charge product
                    # and epsilon are zero (OpenMM DrudeGenerator.postprocessSystem)
                    da = ff.drude[t.atom_types[order[a]]][0]
                    db = ff.drude[t.atom_types[order[b]]][0]
                    qq = ff.coulomb14scale * ff.charge[da] * ff.charge[db]
                    ee = ff.lj14scale * np.sqrt(ff.epsilon[da] * ff.epsilon[db])
                    excluded = (qq == 0.0) and (ee == 0.0)
                if excluded:
                    u[a, b] = u[b, a] = (th[a] + th[b]) / (al[a] * al[b]) ** (1.0 / 6.0)
        names = [tnames[ti] for ti in order]
        if site_names is None:
            site_names = names
        elif names != site_names:
            raise ValueError("heterogeneous residues: the reference layout needs identical residues")
        r_all.append(r); qc_all.append(qc); qs_all.append(qs); al_all.append(al); u_all.append(u)

    return DrudeSystem(
        r_core=np.array(r_all), q_core=np.array(qc_all), q_shell=np.array(qs_all),
        alpha=np.array(al_all), thole_u=np.array(u_all), site_names=site_names,
    )
