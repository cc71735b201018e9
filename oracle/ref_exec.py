"""TEST INFRASTRUCTURE - executes the reference's OWN source for the hot path.

``load_reference()`` imports ``/root/reference/U_pol/polarization.py`` *unmodified, from where
it lies* and returns the module, so that ``Uind``, ``Ucoul``, ``Ucoul_static``, ``Uself``,
``make_Sij`` and ``drudeOpt`` (``polarization.py:27-188``) can be called on real inputs.

The reference's third-party imports are not installable in this image (jax, jaxlib, jaxopt,
optax, openmm: no wheels, no network).  They are replaced by small stand-ins registered in
``sys.modules`` *before* the import:

* ``jax.numpy``  -> a namespace that forwards the handful of functions the hot path uses
  (``exp where isfinite eye sum reshape ravel array zeros inf newaxis linalg.norm``) to
  torch-fp64, so that every arithmetic line of the reference runs as written and
  ``jax.grad`` can be taken with torch's reverse-mode autodiff;
* ``jax.jit``    -> identity;  ``jax.config.update`` -> no-op (fp64 is always on here, which
  is what ``main()`` sets at ``polarization.py:196``);
* ``optax.safe_norm`` -> a restatement of optax 0.2.3's function (norm with value
  ``min_norm`` and zero gradient where the norm is <= ``min_norm``);
* ``jaxopt.BFGS`` -> SciPy BFGS on the reference's closure with the torch gradient, stopping
  on ``||g||_2 <= tol`` (jaxopt's stop rule).  Iterate sequences therefore are NOT the
  reference's; converged energies are comparable.
* ``openmm*``    -> empty stubs (only names; nothing of OpenMM is executed).

Used only by ``tests/golden/make_golden.py`` (to generate committed fixtures) and by CPU tests
that are skipped when ``/root/reference`` is absent (it does not exist on the GPU box).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("UPOL_REFERENCE_ROOT", "/root/reference")
_F64 = torch.float64


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "U_pol", "polarization.py"))


def _as_tensor(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def _make_jnp():
    jnp = types.ModuleType("jax.numpy")
    jnp.newaxis = None
    jnp.inf = float("inf")
    jnp.exp = lambda x: torch.exp(_as_tensor(x))
    jnp.isfinite = torch.isfinite
    jnp.eye = lambda n: torch.eye(int(n), dtype=_F64)
    jnp.sum = lambda x, *a, **k: torch.sum(x, *a, **k)
    jnp.reshape = lambda x, shape: torch.reshape(_as_tensor(x), tuple(shape))
    jnp.ravel = lambda x: torch.reshape(_as_tensor(x), (-1,))
    jnp.array = lambda x: _as_tensor(x).clone()
    jnp.zeros = lambda shape: torch.zeros(shape, dtype=_F64)
    jnp.ones_like = torch.ones_like

    def where(c, a, b):
        if not isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
            a = torch.as_tensor(a, dtype=_F64)
        return torch.where(c, a, b)

    jnp.where = where
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.norm = lambda x, ord=None, axis=None, keepdims=False: torch.linalg.norm(
        x, ord=ord, dim=axis, keepdim=keepdims)
    jnp.linalg = linalg
    return jnp


def _safe_norm(x, min_norm, ord=None, axis=None, keepdims=False):
    # optax 0.2.3 _src/numerics.py safe_norm, restated with torch
    x = _as_tensor(x)
    norm = torch.linalg.norm(x, ord=ord, dim=axis, keepdim=True)
    x = torch.where(norm <= min_norm, torch.ones_like(x), x)
    norm = norm if keepdims else torch.squeeze(norm, dim=axis)
    masked = torch.linalg.norm(x, ord=ord, dim=axis, keepdim=keepdims)
    return torch.where(norm <= min_norm, torch.full_like(masked, float(min_norm)), masked)


class _BFGS:
    """Stand-in for jaxopt.BFGS(fun, tol).run(init_params) -> object with .params"""

    def __init__(self, fun, tol=1e-3, maxiter=500, verbose=False, **_):
        self.fun, self.tol, self.maxiter = fun, tol, maxiter

    def run(self, init_params):
        from scipy.optimize import minimize

        x0 = _as_tensor(init_params).detach().numpy().astype(np.float64)
        shape = x0.shape

        def fg(x):
            t = torch.as_tensor(x.reshape(shape)).clone().requires_grad_(True)
            e = self.fun(t)
            (g,) = torch.autograd.grad(e, t)
            return float(e.detach()), g.numpy().ravel()

        res = minimize(fg, x0.ravel(), jac=True, method="BFGS",
                       options={"gtol": self.tol, "norm": 2, "maxiter": self.maxiter})
        out = types.SimpleNamespace()
        out.params = torch.as_tensor(res.x.reshape(shape))
        out.state = res
        return out


def _install_stubs():
    jnp = _make_jnp()
    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **k: f if f is not None else (lambda g: g)
    jax.config = types.SimpleNamespace(update=lambda *a, **k: None)

    def grad(f, argnums=0):
        def df(*args):
            args = list(args)
            t = _as_tensor(args[argnums]).clone().requires_grad_(True)
            args[argnums] = t
            (g,) = torch.autograd.grad(f(*args), t)
            return g
        return df

    jax.grad = grad
    optax = types.ModuleType("optax")
    optax.safe_norm = _safe_norm
    jaxopt = types.ModuleType("jaxopt")
    jaxopt.BFGS = _BFGS
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp.linalg, "optax": optax, "jaxopt": jaxopt}
    for name, attrs in (
        ("openmm", ["DrudeForce", "NonbondedForce", "DrudeSCFIntegrator", "Platform"]),
        ("openmm.app", ["Simulation", "Topology", "ForceField", "PDBFile", "Modeller"]),
        ("openmm.unit", ["elementary_charge", "picoseconds", "nanometer", "kilojoules_per_mole"]),
    ):
        m = types.ModuleType(name)
        for a in attrs:
            # `picoseconds` is multiplied by a float in a default argument (util.py:286)
            setattr(m, a, 1.0 if name == "openmm.unit" else type(a, (), {}))
        mods[name] = m
    mods["openmm"].app = mods["openmm.app"]
    mods["openmm"].unit = mods["openmm.unit"]
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    return saved


_CACHE = {}


def load_reference():
    """Import the reference's polarization.py (unmodified) under the stand-ins and return it."""
    if "mod" in _CACHE:
        return _CACHE["mod"]
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT}/U_pol/polarization.py not found")
    upol_dir = os.path.join(REFERENCE_ROOT, "U_pol")
    saved = _install_stubs()
    saved_util = sys.modules.pop("util", None)
    sys.path.insert(0, upol_dir)
    old_dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True          # the reference tree is read-only
    try:
        spec = importlib.util.spec_from_file_location("_upol_reference_polarization",
                                                      os.path.join(upol_dir, "polarization.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.set_constants()
    finally:
        sys.dont_write_bytecode = old_dont_write
        sys.path.remove(upol_dir)
        sys.modules.pop("util", None)
        if saved_util is not None:
            sys.modules["util"] = saved_util
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["mod"] = mod
    return mod


def ref_uind_and_grad(Rij, Dij, Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k):
    """Reference ``Uind`` (polarization.py:111-151) and its reverse-mode gradient w.r.t. Dij."""
    ref = load_reference()
    args = [_as_tensor(a) if not np.isscalar(a) else a
            for a in (Rij, Dij, Qi_shell, Qj_shell, Qi_core, Qj_core, u_scale, k)]
    D = args[1].clone().requires_grad_(True)
    args[1] = D
    e = ref.Uind(*args)
    (g,) = torch.autograd.grad(e, D)
    return float(e.detach()), g.numpy()
