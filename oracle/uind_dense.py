"""TEST INFRASTRUCTURE - CPU oracle for the U_ind path.  Not part of the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm
may import this module.  The product (``u_pol_b200``) never does.

Dense torch-fp64 restatement of the reference algorithm, term by term, on the reference's own
dense broadcast layout (``Rij (nmol,nmol,natoms,natoms,3)`` etc.), so that the gradient with
respect to ``Dij`` comes from reverse-mode autodiff exactly as ``jax.grad`` produces it inside
``jaxopt.BFGS`` in the reference (``U_pol/polarization.py:172-179``).

Pinning (see ``tests/test_oracle_goldens.py``):
  * checked against the reference's recorded goldens (``ONE_4PI_EPS0``, imidazole-trimer
    ``U_coul(d=0)``, 312 imidazole-dimer minimised ``U_ind`` values of
    ``U_pol/tests/imidazole/induction_data.csv``), and
  * against outputs of the reference's *own source* (``U_pol/polarization.py``) executed in the
    build container with a torch-backed stand-in for ``jax.numpy`` (``oracle/ref_exec.py``);
    those outputs are committed under ``tests/golden/``.
The minimiser iterate sequence of jaxopt 0.8.3 is NOT pinned (jaxopt is not vendored in the
reference and not installable here); only converged energies are.
"""
from __future__ import annotations

import numpy as np
import torch

# polarization.py:19-25
_PI = 3.14159265358979323846
_E = 1.602176634e-19
_NA = 6.02214076e23
_EPS0 = 1e-6 * 8.8541878128e-12 / (_E * _E * _NA)
COULOMB = 1 / (4 * _PI * _EPS0)

_F64 = torch.float64


def _t(x):
    if isinstance(x, torch.Tensor):
        return x.to(_F64)
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def norm0(x):
    """optax.safe_norm(x, 0.0, axis=-1): value 0 and gradient 0 at the origin
    (used at polarization.py:30,36,47)."""
    sq = (x * x).sum(-1)
    pos = sq > 0
    return torch.where(pos, torch.sqrt(torch.where(pos, sq, torch.ones_like(sq))), torch.zeros_like(sq))


def inv_len(x):
    """1/|x| with the reference's convention |x| = 0 -> +inf -> term 0 (polarization.py:33-37)."""
    n = norm0(x)
    return 1.0 / torch.where(n == 0, torch.full_like(n, float("inf")), n)


def thole_S(x, u):
    """S = 1 - (1 + u r / 2) exp(-u r), r = |x|   (polarization.py:27-31)."""
    r = norm0(x)
    return 1.0 - (1.0 + 0.5 * r * u) * torch.exp(-u * r)


def finite_sum(x):
    """polarization.py:39-42."""
    return torch.where(torch.isfinite(x), x, torch.zeros_like(x)).sum()


def spring_energy(D, k):
    """polarization.py:44-48."""
    return 0.5 * (k * norm0(D) ** 2).sum()


def _mol_masks(nmol, natoms):
    same_mol = torch.eye(nmol, dtype=_F64)[:, :, None, None]
    same_atom = torch.eye(natoms, dtype=_F64)[None, None, :, :]
    return same_mol, same_atom


def coulomb_static(R, qis, qjs, qic, qjc):
    """polarization.py:50-61: total-charge Coulomb energy between different molecules."""
    nmol = R.shape[0]
    same_mol, _ = _mol_masks(nmol, R.shape[2])
    e = (qic + qis) * (qjc + qjs) * inv_len(R) * (1 - same_mol)
    return COULOMB * 0.5 * finite_sum(e)


def coulomb_total(R, D, qis, qjs, qic, qjc, u):
    """polarization.py:63-109."""
    nmol, natoms = R.shape[0], R.shape[2]
    di = D[:, None, :, None, :]
    dj = D[None, :, None, :, :]
    x_cc, x_sc, x_cs, x_ss = R, R + di, R - dj, R + di - dj          # :70-73
    f_cc, f_sc, f_cs, f_ss = inv_len(x_cc), inv_len(x_sc), inv_len(x_cs), inv_len(x_ss)
    s_cc, s_sc, s_cs, s_ss = (thole_S(x, u) for x in (x_cc, x_sc, x_cs, x_ss))   # :76-79
    same_mol, same_atom = _mol_masks(nmol, natoms)
    inter = (qic * qjc * f_cc + qis * qjc * f_sc + qic * qjs * f_cs + qis * qjs * f_ss) * (1 - same_mol)  # :82-91
    intra = (
        s_cc * (-qis) * (-qjs) * f_cc + s_sc * qis * (-qjs) * f_sc
        + s_cs * (-qis) * qjs * f_cs + s_ss * qis * qjs * f_ss
    ) * same_mol * (1 - same_atom)                                   # :94-105
    return COULOMB * 0.5 * finite_sum(inter + intra)                 # :107-109


def u_ind(R, D, qis, qjs, qic, qjc, u, k):
    """polarization.py:111-151; every argument in the reference's dense layout."""
    R, qis, qjs, qic, qjc, k = (_t(a) for a in (R, qis, qjs, qic, qjc, k))
    u = _t(u) if not np.isscalar(u) else torch.tensor(float(u), dtype=_F64)
    D = _t(D)
    nmol, natoms = R.shape[0], R.shape[2]
    D = D.reshape(nmol, natoms, 3)
    return (coulomb_total(R, D, qis, qjs, qic, qjc, u) - coulomb_static(R, qis, qjs, qic, qjc)) + spring_energy(D, k)


def u_ind_and_grad(R, D, qis, qjs, qic, qjc, u, k):
    """(U_ind, dU/dD) with reverse-mode AD, the stand-in for jax.value_and_grad."""
    D = _t(D).clone().requires_grad_(True)
    e = u_ind(R, D, qis, qjs, qic, qjc, u, k)
    (g,) = torch.autograd.grad(e, D)
    return float(e.detach()), g.detach().numpy().reshape(tuple(D.shape))


def minimise(R, D0, qis, qjs, qic, qjc, u, k, gtol=1e-4, maxiter=500):
    """Stand-in for ``jaxopt.BFGS(fun=Uind_min, tol=gtol).run(D0)`` (polarization.py:178-179):
    SciPy BFGS stopping on the 2-norm of the gradient.  Returns (d_opt flat, U_ind, info)."""
    from scipy.optimize import minimize

    shape = np.asarray(D0).shape
    x0 = np.asarray(D0, dtype=np.float64).ravel()

    def fg(x):
        e, g = u_ind_and_grad(R, x, qis, qjs, qic, qjc, u, k)
        return e, g.ravel()

    res = minimize(fg, x0, jac=True, method="BFGS", options={"gtol": gtol, "norm": 2, "maxiter": maxiter})
    return res.x.reshape(-1), float(res.fun), {"nit": int(res.nit), "nfev": int(res.nfev),
                                                "gnorm": float(np.linalg.norm(res.jac))}


def newton_minimum(R, D0, qis, qjs, qic, qjc, u, k, gtol=1e-10, maxiter=50):
    """Ground-truth minimum for SMALL systems: Newton iterations on the Drude dof with the exact
    (autodiff) Hessian.  U_ind is near-quadratic in d, so this reaches ||g||_2 ~ 1e-11 where the
    Wolfe line search of BFGS stalls on round-off.  Returns (d_opt, U, gnorm)."""
    R, qis, qjs, qic, qjc, k = (_t(a) for a in (R, qis, qjs, qic, qjc, k))
    ut = _t(u) if not np.isscalar(u) else torch.tensor(float(u), dtype=_F64)
    nmol, natoms = R.shape[0], R.shape[2]
    act = (qis.reshape(nmol, natoms) != 0)                       # polarisable sites only
    idx = act.reshape(-1).nonzero().reshape(-1)
    x = _t(D0).reshape(nmol * natoms, 3)[idx].clone()

    def f(xa):
        D = torch.zeros(nmol * natoms, 3, dtype=_F64).index_put((idx,), xa)
        return u_ind(R, D, qis, qjs, qic, qjc, ut, k)

    gn = float("inf")
    for _ in range(maxiter):
        xa = x.clone().requires_grad_(True)
        e = f(xa)
        (g,) = torch.autograd.grad(e, xa, create_graph=True)
        gn = float(g.detach().norm())
        if gn <= gtol:
            break
        H = torch.stack([torch.autograd.grad(gi, xa, retain_graph=True)[0].reshape(-1) for gi in g.reshape(-1)])
        step = torch.linalg.solve(H, g.detach().reshape(-1)).reshape(x.shape)
        x = x - step
    D = torch.zeros(nmol * natoms, 3, dtype=_F64).index_put((idx,), x)
    return D.reshape(-1).numpy(), float(f(x)), gn


# --- dense-array construction from compact site arrays (util.py:73-84,138-147,264-279) ----

def dense_inputs(r_core, q_core, q_shell, k, thole_u):
    """Compact (nmol,natoms,..) arrays -> the reference's dense broadcast tensors."""
    r = np.asarray(r_core, dtype=np.float64)
    nmol = r.shape[0]
    R = r[None, :, None, :, :] - r[:, None, :, None, :]
    qs = np.asarray(q_shell, dtype=np.float64)
    qc = np.asarray(q_core, dtype=np.float64)
    qis, qjs = qs[:, None, :, None], qs[None, :, None, :]
    qic, qjc = qc[:, None, :, None], qc[None, :, None, :]
    th = np.asarray(thole_u, dtype=np.float64)
    if np.any(th != 0):
        u = th[None, ...] * np.eye(nmol)[:, :, None, None]
    else:
        u = 0.0
    return R, qis, qjs, qic, qjc, u, np.asarray(k, dtype=np.float64)
