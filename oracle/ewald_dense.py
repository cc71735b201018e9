"""TEST INFRASTRUCTURE - periodic (Ewald) restatement of U_ind for SURVEY 8f "next" #3 (second half).

PARITY UNPINNED AGAINST THE REFERENCE: shehan807/U_pol is gas phase only (open boundaries,
``benchmarks/OpenMM/water/calc_energy.py:80``), so there is no reference value for a periodic box.
What is pinned instead (``tests/test_oracle_ewald.py``): the Madelung constants of NaCl and CsCl,
independence of the splitting parameter kappa, and the open-boundary limit (box -> infinity) against
the pinned open-boundary oracle ``uind_dense`` (itself checked against the reference's goldens).

Definition (the natural periodic extension of polarization.py:63-151).  The reference sums
q_a q_b / |x_a - x_b| over charges of DIFFERENT molecules (mask 1 - delta_mn, :90-91) and adds the
Thole-screened intra-molecular shell terms (:94-105).  Under periodic boundaries every charge also
sees all periodic images: pairs of different molecules in every image cell, pairs of the same molecule
(and a charge with itself) in the image cells n != 0 only.  With tin-foil boundary conditions that
lattice sum is the Ewald sum with intra-molecular exclusions:

    E = K [ sum_{a<b, other molecule} q_a q_b erfc(kappa r_ab)/r_ab              (real space, all images)
          + (2 pi / V) sum_{k != 0} exp(-k^2 / 4 kappa^2) / k^2 |S(k)|^2          (reciprocal space)
          - kappa / sqrt(pi) sum_a q_a^2                                          (self)
          - sum_{a<b, same molecule} q_a q_b erf(kappa r_ab)/r_ab ]               (excluded pairs, n = 0)

    U_ind = [E(cores at r, shells at r - d) - E(cores at r, shells at r)] + Thole intra + springs,

the last two exactly as in the open-boundary case (they act inside one molecule).  Units nm, e,
kJ/mol; orthorhombic box.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import uind_dense
from .uind_dense import COULOMB

_F64 = torch.float64


def k_vectors(box, kappa, nmax=None, tol_exponent=41.5):
    """Half space of reciprocal vectors with exp(-k^2/4kappa^2) above exp(-tol_exponent) (or inside the
    integer cube |n| <= nmax when given).  Returns (n (M,3) int64, k (M,3), A (M,)) with
    A = exp(-k^2/4kappa^2)/k^2; the half space is n_x > 0, or n_x = 0 and n_y > 0, or n_x = n_y = 0 and
    n_z > 0, so every pair (k, -k) appears once."""
    box = np.asarray(box, dtype=np.float64)
    kmax = 2.0 * kappa * math.sqrt(tol_exponent)
    lim = [int(math.floor(kmax * L / (2 * math.pi))) for L in box] if nmax is None else [int(nmax)] * 3
    rng = [np.arange(-m, m + 1) for m in lim]
    n = np.stack(np.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3)
    half = (n[:, 0] > 0) | ((n[:, 0] == 0) & (n[:, 1] > 0)) | ((n[:, 0] == 0) & (n[:, 1] == 0) & (n[:, 2] > 0))
    n = n[half]
    k = 2 * math.pi * n / box
    k2 = (k * k).sum(1)
    keep = k2 <= kmax * kmax if nmax is None else np.ones(len(n), bool)
    n, k, k2 = n[keep], k[keep], k2[keep]
    return n.astype(np.int64), k, np.exp(-k2 / (4 * kappa * kappa)) / k2


def _erf_over_r(r, kappa):
    """erf(kappa r)/r with its r -> 0 limit 2 kappa/sqrt(pi)."""
    small = r < 1e-9
    rs = torch.where(small, torch.ones_like(r), r)
    return torch.where(small, torch.full_like(r, 2 * kappa / math.sqrt(math.pi)), torch.erf(kappa * rs) / rs)


def ewald_energy(X, Q, mol, box, kappa, nmax=None, images=1, kchunk=4096):
    """Ewald energy (kJ/mol) of point charges Q at X (M,3) in an orthorhombic box; pairs with equal
    ``mol`` are excluded in the home cell (n = 0) but interact through their images.  Real space runs
    over the (2*images+1)^3 nearest cells, so it does not rely on the minimum-image shortcut."""
    X = torch.as_tensor(X, dtype=_F64)
    Q = torch.as_tensor(np.asarray(Q, dtype=np.float64))
    mol = torch.as_tensor(np.asarray(mol, dtype=np.int64))
    boxt = torch.as_tensor(np.asarray(box, dtype=np.float64))
    M = X.shape[0]
    dx = X[:, None, :] - X[None, :, :]
    same = mol[:, None] == mol[None, :]
    qq = Q[:, None] * Q[None, :]
    e_real = torch.zeros((), dtype=_F64)
    cells = range(-images, images + 1)
    for nx in cells:
        for ny in cells:
            for nz in cells:
                shift = torch.tensor([nx, ny, nz], dtype=_F64) * boxt
                # home cell: excluded pairs are masked below; keep their distance away from 0 so that a shell
                # sitting exactly on its core (d = 0) does not put inf * 0 into the backward pass
                rr = torch.sqrt(((dx + shift) ** 2).sum(-1) + (same.to(_F64) if (nx, ny, nz) == (0, 0, 0) else 0.0))
                term = qq * torch.erfc(kappa * rr) / rr
                if (nx, ny, nz) == (0, 0, 0):
                    term = torch.where(same, torch.zeros_like(term), term)
                e_real = e_real + 0.5 * term.sum()
    # excluded pairs of the home cell are still inside the reciprocal sum: take them out again
    r0 = torch.sqrt((dx ** 2).sum(-1) + 1e-300)
    offdiag = same & ~torch.eye(M, dtype=torch.bool)
    e_excl = 0.5 * torch.where(offdiag, qq * _erf_over_r(r0, kappa), torch.zeros_like(qq)).sum()
    e_self = kappa / math.sqrt(math.pi) * (Q * Q).sum()
    _, kv, A = k_vectors(box, kappa, nmax)
    vol = float(np.prod(np.asarray(box, dtype=np.float64)))
    e_rec = torch.zeros((), dtype=_F64)
    for c in range(0, len(kv), kchunk):
        kk = torch.as_tensor(kv[c:c + kchunk])
        ph = X @ kk.T                                            # (M, nk)
        sre = (Q[:, None] * torch.cos(ph)).sum(0)
        sim = (Q[:, None] * torch.sin(ph)).sum(0)
        e_rec = e_rec + (torch.as_tensor(A[c:c + kchunk]) * (sre * sre + sim * sim)).sum()
    e_rec = e_rec * (2.0 * (2 * math.pi / vol))                  # half space -> factor 2
    return COULOMB * (e_real + e_rec - e_self - e_excl)


def _thole_intra(R, D, qis, qjs, u):
    """The intra-molecular part of polarization.py:94-105 alone (home cell, Thole screened)."""
    nmol, natoms = R.shape[0], R.shape[2]
    di = D[:, None, :, None, :]
    dj = D[None, :, None, :, :]
    xs = (R, R + di, R - dj, R + di - dj)
    f = [uind_dense.inv_len(x) for x in xs]
    s = [uind_dense.thole_S(x, u) for x in xs]
    same_mol, same_atom = uind_dense._mol_masks(nmol, natoms)
    intra = (s[0] * qis * qjs * f[0] - s[1] * qis * qjs * f[1] - s[2] * qis * qjs * f[2] + s[3] * qis * qjs * f[3])
    return COULOMB * 0.5 * uind_dense.finite_sum(intra * same_mol * (1 - same_atom))


def u_ind_pbc(r_core, D, q_core, q_shell, k, thole_u, box, kappa, nmax=None, images=1):
    """Periodic U_ind (torch scalar; differentiable in D).  Compact inputs: r_core, D (nmol,natoms,3);
    q_core, q_shell, k (nmol,natoms); thole_u (nmol,natoms,natoms)."""
    r = r_core if isinstance(r_core, torch.Tensor) else torch.as_tensor(np.asarray(r_core, dtype=np.float64))
    D = D if isinstance(D, torch.Tensor) else torch.as_tensor(np.asarray(D, dtype=np.float64))
    nmol, natoms = r.shape[0], r.shape[1]
    D = D.reshape(nmol, natoms, 3)
    qc = np.asarray(q_core, dtype=np.float64).reshape(-1)
    qs = np.asarray(q_shell, dtype=np.float64).reshape(-1)
    pol = np.nonzero(qs != 0)[0]
    mol = np.repeat(np.arange(nmol), natoms)
    rf = r.reshape(-1, 3)
    Q = np.concatenate([qc, qs[pol]])
    M = np.concatenate([mol, mol[pol]])
    pol_t = torch.as_tensor(pol)
    X_d = torch.cat([rf, rf[pol_t] - D.reshape(-1, 3)[pol_t]], 0)
    X_0 = torch.cat([rf, rf[pol_t]], 0)
    e_d = ewald_energy(X_d, Q, M, box, kappa, nmax, images)
    e_0 = ewald_energy(X_0, Q, M, box, kappa, nmax, images)
    _, qis, qjs, _, _, u, kk = uind_dense.dense_inputs(r.detach().numpy(), q_core, q_shell, k, thole_u)
    R = r[None, :, None, :, :] - r[:, None, :, None, :]            # inside the graph: differentiable in r as well
    qis, qjs, kk = (uind_dense._t(a) for a in (qis, qjs, kk))
    u = uind_dense._t(u) if not np.isscalar(u) else torch.tensor(float(u), dtype=_F64)
    return (e_d - e_0) + _thole_intra(R, D, qis, qjs, u) + uind_dense.spring_energy(D, kk)


def u_ind_pbc_and_grad(r_core, D, q_core, q_shell, k, thole_u, box, kappa, nmax=None, images=1):
    Dt = torch.as_tensor(np.asarray(D, dtype=np.float64)).clone().requires_grad_(True)
    e = u_ind_pbc(r_core, Dt, q_core, q_shell, k, thole_u, box, kappa, nmax, images)
    (g,) = torch.autograd.grad(e, Dt)
    return float(e.detach()), g.detach().numpy().reshape(np.asarray(D).shape)


def u_ind_pbc_core_grad(r_core, D, q_core, q_shell, k, thole_u, box, kappa, nmax=None, images=1):
    """dU_ind/dr_core at fixed displacements (the shells move with their cores)."""
    rt = torch.as_tensor(np.asarray(r_core, dtype=np.float64)).clone().requires_grad_(True)
    e = u_ind_pbc(rt, D, q_core, q_shell, k, thole_u, box, kappa, nmax, images)
    (g,) = torch.autograd.grad(e, rt)
    return g.detach().numpy()


def default_kappa(box):
    """Splitting parameter for which the home-cell (minimum image) real-space sum is complete to
    ~1e-19: erfc(kappa L_min / 2) = erfc(6.5)."""
    return 13.0 / float(np.min(np.asarray(box, dtype=np.float64)))
