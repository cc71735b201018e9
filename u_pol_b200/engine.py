"""Scalable host API over the C ABI: compact O(N) inputs, device-resident state.

``UindSystem`` is the added scalable entry of SURVEY 8b: same semantics as the reference's
``Uind`` / ``drudeOpt`` (U_pol/polarization.py:111-188) without the dense
(nmol,nmol,natoms,natoms[,3]) tensors of util.py:78-81,271-274.  PyTorch is used for device
memory and streams only; all arithmetic happens in libupol_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _np(a, dtype):
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.UpolError("u_pol_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def thole_pairs_from_blocks(thole_u, natoms):
    """(nmol, natoms, natoms) symmetric u blocks (tholeMatrix of util.py:219-247) -> unordered
    site-pair list (a, b, u) with a < b."""
    th = np.asarray(thole_u, dtype=np.float64)
    m, a, b = np.nonzero(np.triu(th, 1))
    return ((m * natoms + a).astype(np.int32), (m * natoms + b).astype(np.int32), th[m, a, b].copy())


class UindSystem:
    """One geometry + parameter set resident in HBM on ``device``.

    r_core (N,3) nm; q_core, q_shell, k (N,); mol_id (N,) non-decreasing; screened pairs as
    (pair_a, pair_b, pair_u) site-index lists.
    """

    def __init__(self, r_core, q_core, q_shell, k, mol_id, pair_a=None, pair_b=None, pair_u=None, device=None):
        _require_cuda()
        if device is None:
            idx = torch.cuda.current_device()
        elif isinstance(device, int):
            idx = device
        else:
            dev = torch.device(device)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        r = _np(r_core, np.float64).reshape(-1, 3)
        self.nsite = r.shape[0]
        qc, qs, kk = (_np(x, np.float64).reshape(-1) for x in (q_core, q_shell, k))
        mol = _np(mol_id, np.int32).reshape(-1)
        if not (qc.size == qs.size == kk.size == mol.size == self.nsite):
            raise ValueError("r_core, q_core, q_shell, k, mol_id disagree on the number of sites")
        if pair_a is None:
            pa = pb = np.zeros(0, np.int32)
            pu = np.zeros(0, np.float64)
        else:
            pa, pb, pu = _np(pair_a, np.int32), _np(pair_b, np.int32), _np(pair_u, np.float64)
        self._keep = (r, qc, qs, kk, mol, pa, pb, pu)
        h = C.c_void_p()
        L.check(L.lib().upol_system_create(
            C.byref(h), self.device.index, self.nsite, r.ctypes.data, qc.ctypes.data, qs.ctypes.data,
            kk.ctypes.data, mol.ctypes.data, pa.size, pa.ctypes.data, pb.ctypes.data, pu.ctypes.data),
            "upol_system_create")
        self._h = h
        self.ndrude = int(L.lib().upol_system_ndrude(h))
        self.nmol = int(np.unique(mol).size)
        self.rank, self.nranks = 0, 1

    # -- construction helpers -------------------------------------------------------------
    @classmethod
    def from_compact(cls, r_core, q_core, q_shell, k, thole_u=None, device=None):
        """Reference-shaped compact arrays: r_core (nmol,natoms,3), q_* and k (nmol,natoms),
        thole_u (nmol,natoms,natoms) or None."""
        r = _np(r_core, np.float64)
        nmol, natoms = r.shape[0], r.shape[1]
        mol = np.repeat(np.arange(nmol, dtype=np.int32), natoms)
        pa = pb = pu = None
        if thole_u is not None and np.any(np.asarray(thole_u) != 0):
            pa, pb, pu = thole_pairs_from_blocks(thole_u, natoms)
        return cls(r.reshape(-1, 3), _np(q_core, np.float64).reshape(-1), _np(q_shell, np.float64).reshape(-1),
                   _np(k, np.float64).reshape(-1), mol, pa, pb, pu, device=device)

    @classmethod
    def from_drude_system(cls, system, device=None):
        return cls.from_compact(system.r_core, system.q_core, system.q_shell, system.k, system.thole_u, device=device)

    def close(self):
        if getattr(self, "_h", None):
            L.lib().upol_system_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev_f64(self, x, n):
        t = torch.as_tensor(x, dtype=torch.float64) if not isinstance(x, torch.Tensor) else x.to(torch.float64)
        t = t.to(self.device).reshape(-1).contiguous()
        if t.numel() != n:
            raise ValueError(f"expected {n} values, got {t.numel()}")
        return t

    def interactions(self, static_part=False):
        return int(L.lib().upol_system_interactions(self._h, int(static_part)))

    def set_positions(self, r_core):
        if isinstance(r_core, torch.Tensor) and r_core.is_cuda:
            t = self._dev_f64(r_core, 3 * self.nsite)
            L.check(L.lib().upol_system_set_positions_dev(self._h, t.data_ptr(), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()
        else:
            r = _np(r_core, np.float64).reshape(-1)
            L.check(L.lib().upol_system_set_positions(self._h, r.ctypes.data, self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    # -- hot path -------------------------------------------------------------------------
    def energy_grad(self, d, precision="fp64", want_grad=True, want_energy=True):
        """(energy vector [8] on device or None, grad (N,3) on device or None).  Asynchronous.
        ``precision`` = "fp64" or "mixed" (fp32 pair arithmetic in cancellation-free difference form with fp64
        accumulation, for the energy and the gradient alike; 1e-5 relative)."""
        prec = {"fp64": L.FP64, "mixed": L.MIXED}[precision]
        with torch.cuda.device(self.device):
            dd = self._dev_f64(d, 3 * self.nsite)
            e = torch.empty(L.E_SIZE, dtype=torch.float64, device=self.device) if want_energy else None
            g = torch.empty(self.nsite, 3, dtype=torch.float64, device=self.device) if want_grad else None
            L.check(L.lib().upol_uind_energy_grad_dev(self._h, dd.data_ptr(), prec,
                                                       e.data_ptr() if want_energy else None,
                                                       g.data_ptr() if want_grad else None, self._stream()),
                    "upol_uind_energy_grad_dev")
        return e, g

    def energy_grad_host(self, d, precision="fp64", want_grad=True):
        """Host-buffer variant (numpy in, numpy out; copies inside the call)."""
        prec = {"fp64": L.FP64, "mixed": L.MIXED}[precision]
        dd = _np(d, np.float64).reshape(-1)
        e = np.empty(L.E_SIZE, np.float64)
        g = np.empty(3 * self.nsite, np.float64) if want_grad else None
        L.check(L.lib().upol_uind_energy_grad_host(self._h, dd.ctypes.data, prec, e.ctypes.data,
                                                    g.ctypes.data if want_grad else None),
                "upol_uind_energy_grad_host")
        return e, (g.reshape(self.nsite, 3) if want_grad else None)

    def energy_grad_host_sharded(self, d, r_core=None, precision="fp64", energy=None, grad=None):
        """Row-sharded host path (``upol_uind_energy_grad_host_sharded``): only this rank's site slice of
        ``r_core`` (optional new geometry) and ``d`` goes up, only its slice of the gradient comes down.
        ``energy`` (8,) / ``grad`` (N,3) may be preallocated (page-locked) numpy arrays; returns
        (energy, grad, (site_lo, site_hi)) - grad holds this rank's rows only, the energy is the total."""
        prec = {"fp64": L.FP64, "mixed": L.MIXED}[precision]
        dd = d if isinstance(d, np.ndarray) and d.dtype == np.float64 and d.flags.c_contiguous else _np(d, np.float64)
        rr = None
        if r_core is not None:
            rr = r_core if isinstance(r_core, np.ndarray) and r_core.dtype == np.float64 and r_core.flags.c_contiguous \
                else _np(r_core, np.float64)
        e = np.empty(L.E_SIZE, np.float64) if energy is None else energy
        g = np.zeros(3 * self.nsite, np.float64) if grad is None else grad
        lo, hi = C.c_int64(), C.c_int64()
        L.check(L.lib().upol_uind_energy_grad_host_sharded(
            self._h, rr.ctypes.data if rr is not None else None, dd.ctypes.data, prec, e.ctypes.data, g.ctypes.data,
            C.byref(lo), C.byref(hi)), "upol_uind_energy_grad_host_sharded")
        return e, g.reshape(self.nsite, 3), (lo.value, hi.value)

    def core_gradient(self, d):
        """dU_ind/dr_core at fixed displacements, (N,3) on device (force on the cores = minus this).
        Not in the reference."""
        with torch.cuda.device(self.device):
            dd = self._dev_f64(d, 3 * self.nsite)
            out = torch.empty(self.nsite, 3, dtype=torch.float64, device=self.device)
            L.check(L.lib().upol_uind_core_gradient_dev(self._h, dd.data_ptr(), out.data_ptr(), self._stream()),
                    "upol_uind_core_gradient_dev")
        return out

    def coulomb_static(self):
        """(K sum qc qc / R, K sum Q Q / R) over inter-molecular pairs, on device."""
        with torch.cuda.device(self.device):
            out = torch.empty(2, dtype=torch.float64, device=self.device)
            L.check(L.lib().upol_coulomb_static_dev(self._h, out.data_ptr(), self._stream()))
        return out

    def minimize(self, d0=None, tol=1e-4, maxiter=500, method="lbfgs", precision="fp64", history=8, check_every=4,
                 strict=False):
        """Minimise U_ind over the Drude displacements on the device.
        Returns (d_opt (N,3) device tensor, result dict).  A lost peer (UPOL_FLAG_COMM) always raises; with
        ``strict`` a non-finite gradient raises too and an unconverged run warns (what ``drudeOpt`` does) -
        otherwise those flags are only reported in ``info['flags']``."""
        with torch.cuda.device(self.device):
            d = torch.zeros(3 * self.nsite, dtype=torch.float64, device=self.device) if d0 is None \
                else self._dev_f64(d0, 3 * self.nsite).clone()
            o = L.MinOptions()
            L.lib().upol_min_options_default(C.byref(o))
            o.tol, o.maxiter = float(tol), int(maxiter)
            o.method = {"lbfgs": L.LBFGS, "bfgs": L.LBFGS, "cg": L.CG}[method.lower()]
            o.precision = {"fp64": L.FP64, "mixed": L.MIXED}[precision]
            o.history, o.check_every = int(history), int(check_every)
            res = L.MinResult()
            rc = L.check(L.lib().upol_minimize_dev(self._h, d.data_ptr(), C.byref(o), C.byref(res), self._stream()),
                         "upol_minimize_dev")
        info = {"energy": np.array(res.energy[:]), "U_ind": res.energy[L.E_UIND], "gnorm": res.gnorm,
                "iterations": res.iterations, "evaluations": res.evaluations, "flags": res.flags, "status": rc}
        if strict:
            L.check_min_flags(res.flags, "UindSystem.minimize")
        elif res.flags & L.FLAG_COMM:
            raise L.UpolError("UindSystem.minimize: a peer rank did not deliver within the time-out (UPOL_FLAG_COMM)")
        return d.reshape(self.nsite, 3), info

    def invalidate_static(self):
        """Forget the cached d-independent part (as a new geometry would): the next energy
        evaluation recomputes it.  Used by bench.py so that a timed step does all the work."""
        L.check(L.lib().upol_system_set_row_range(self._h, *self._mol_range()))

    def _mol_range(self):
        return getattr(self, "_mrange", (0, self.nmol))

    def set_kernel_timing(self, enable=True):
        L.check(L.lib().upol_system_set_kernel_timing(self._h, int(enable)))

    def launch_count(self):
        """Kernels of the library launched for this system so far."""
        return int(L.lib().upol_system_launch_count(self._h))

    def kernel_time(self):
        """(summed ms, launches) of the main pair kernel since the last call."""
        ms, n = C.c_double(), C.c_int32()
        L.check(L.lib().upol_system_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- multi-GPU ------------------------------------------------------------------------
    def attach_process_group(self, group=None, p2p=True):
        """Row-shard this system over the ranks of a torch.distributed process group: creates the
        library's own NCCL communicator (id shipped with torch.distributed) and sets this rank's
        molecule block.  With ``p2p`` the minimiser loop then runs on peer-memory stores
        (``enable_p2p``) and NCCL is used only for the energy/gradient entry points."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_ubyte * L.COMM_ID_BYTES)()
        if rank == 0:
            L.check(L.lib().upol_comm_unique_id(buf), "upol_comm_unique_id")
        ids = [bytes(buf)]
        dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = (C.c_ubyte * L.COMM_ID_BYTES).from_buffer_copy(ids[0])
        with torch.cuda.device(self.device):
            L.check(L.lib().upol_system_attach_comm(self._h, raw, rank, world), "upol_system_attach_comm")
        self.rank, self.nranks = rank, world
        self._mrange = shard_bounds(self.nmol, world, rank)
        if p2p and world > 1:
            self.enable_p2p(group)

    def enable_p2p(self, group=None):
        """Map every rank's displacement vector / mailbox / flags into every other rank (CUDA IPC over
        NVLink) so that a minimiser iteration exchanges data with plain peer stores instead of NCCL."""
        import torch.distributed as dist

        blob = (C.c_ubyte * L.P2P_EXPORT_BYTES)()
        with torch.cuda.device(self.device):
            ok = L.lib().upol_system_p2p_export(self._h, blob) == 0
        got = [None] * dist.get_world_size(group)
        dist.all_gather_object(got, (ok, bytes(blob)), group=group)
        ok = all(g[0] for g in got)
        if ok:
            allb = (C.c_ubyte * (L.P2P_EXPORT_BYTES * len(got))).from_buffer_copy(b"".join(g[1] for g in got))
            with torch.cuda.device(self.device):
                ok = L.lib().upol_system_p2p_attach(self._h, allb) == 0
        agree = [None] * dist.get_world_size(group)
        dist.all_gather_object(agree, ok, group=group)
        if not all(agree):                      # some rank could not map its peers: everyone stays on NCCL
            with torch.cuda.device(self.device):
                L.lib().upol_system_p2p_disable(self._h)
        self.p2p = all(agree)
        dist.barrier(group)
        return self.p2p

    def set_periodic(self, box, kappa=None, nmax=0):
        """Periodic boundaries (an extension: the reference is gas phase only).  ``box`` = 3 edge lengths
        in nm of an orthorhombic cell, or None to return to open boundaries.  U_ind, its gradient and the
        minimiser then refer to the periodic system (direct Ewald sum, tin-foil; definition and checker:
        ``oracle/ewald_dense.py``).  ``kappa`` defaults to 13 / min(box); ``nmax`` > 0 replaces the default
        reciprocal cut-off exp(-k^2/4kappa^2) >= exp(-41.5) by the cube |n| <= nmax.  fp64; works with row
        shards (``attach_process_group``) and ``core_gradient`` like the open-boundary path."""
        if box is None:
            L.check(L.lib().upol_system_set_periodic(self._h, None, 0.0, 0), "upol_system_set_periodic")
            return None
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64).reshape(3))
        L.check(L.lib().upol_system_set_periodic(self._h, b.ctypes.data, float(kappa or 0.0), int(nmax)),
                "upol_system_set_periodic")
        return self.periodic_info()

    def periodic_info(self):
        """{'box', 'kappa', 'n_kvectors'} of the periodic set-up, or None for open boundaries."""
        b = np.zeros(3)
        kap, nk = C.c_double(), C.c_int32()
        if L.lib().upol_system_periodic_info(self._h, b.ctypes.data, C.byref(kap), C.byref(nk)) != 0:
            return None
        return {"box": b, "kappa": kap.value, "n_kvectors": nk.value}

    def set_row_range(self, mol_lo, mol_hi):
        L.check(L.lib().upol_system_set_row_range(self._h, int(mol_lo), int(mol_hi)))
        self._mrange = (int(mol_lo), int(mol_hi))


def shard_bounds(nmol, nranks, rank):
    """Molecule block of ``rank`` (same rule as csrc/comm.cu::shard_bounds)."""
    return nmol * rank // nranks, nmol * (rank + 1) // nranks


def measure_fma_peak(fp64=True, device=0):
    """FLOP/s of a register-resident FMA chain: fp64=True DFMA, False FFMA, "f32x2" packed FFMA2."""
    _require_cuda()
    v = C.c_double()
    mode = 2 if fp64 == "f32x2" else int(bool(fp64))
    L.check(L.lib().upol_measure_fma_peak(int(device), mode, C.byref(v)))
    return v.value


def batched_minimize(r_batch, q_core, q_shell, k, thole_u=None, d0=None, tol=1e-4, maxiter=500, device=None):
    """Minimise U_ind for a batch of independent small systems sharing one topology, in one launch
    (one warp per system; replaces the subprocess-per-dimer loop of wrapper.py:265-316).

    r_batch (B, nmol, natoms, 3) nm; q_core, q_shell, k (nmol, natoms); thole_u (nmol, natoms, natoms).
    Returns (d_opt (B, nmol*natoms, 3), energy (B, 8), iterations (B,)) as CUDA tensors; energy[:, 6]
    holds the UPOL_FLAG_* bits of each system."""
    _require_cuda()
    idx = torch.cuda.current_device() if device is None else int(torch.device(device).index or 0) if not isinstance(device, int) else device
    dev = torch.device("cuda", idx)
    r = torch.as_tensor(r_batch, dtype=torch.float64) if not isinstance(r_batch, torch.Tensor) else r_batch.to(torch.float64)
    if r.dim() != 4 or r.shape[-1] != 3:
        raise ValueError("r_batch must be (B, nmol, natoms, 3)")
    B, nmol, natoms = int(r.shape[0]), int(r.shape[1]), int(r.shape[2])
    nsite = nmol * natoms
    qc, qs, kk = (_np(a, np.float64).reshape(-1) for a in (q_core, q_shell, k))
    if not (qc.size == qs.size == kk.size == nsite):
        raise ValueError("q_core, q_shell, k must be (nmol, natoms)")
    mol = np.repeat(np.arange(nmol, dtype=np.int32), natoms)
    if thole_u is not None and np.any(np.asarray(thole_u) != 0):
        pa, pb, pu = thole_pairs_from_blocks(thole_u, natoms)
    else:
        pa = pb = np.zeros(0, np.int32)
        pu = np.zeros(0, np.float64)
    with torch.cuda.device(dev):
        rd = r.to(dev).reshape(B, nsite * 3).contiguous()
        d = torch.zeros(B, nsite * 3, dtype=torch.float64, device=dev) if d0 is None else \
            torch.as_tensor(d0, dtype=torch.float64).to(dev).reshape(B, nsite * 3).clone()
        e = torch.zeros(B, L.E_SIZE, dtype=torch.float64, device=dev)
        it = torch.zeros(B, dtype=torch.int32, device=dev)
        L.check(L.lib().upol_batched_minimize_dev(
            idx, B, nsite, rd.data_ptr(), d.data_ptr(), qc.ctypes.data, qs.ctypes.data, kk.ctypes.data,
            mol.ctypes.data, pa.size, pa.ctypes.data, pb.ctypes.data, pu.ctypes.data, float(tol), int(maxiter),
            e.data_ptr(), it.data_ptr(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
            "upol_batched_minimize_dev")
    return d.reshape(B, nsite, 3), e, it


def batched_minimize_replicas(r_batch, q_core, q_shell, k, thole_u=None, tol=1e-4, maxiter=500, group=None):
    """BASELINE config 5 on several GPUs: independent systems need no data-path collective ("replicas
    only") - every rank minimises its contiguous share of the batch on its own GPU and the
    results are all-gathered for convenience.  Returns (d_opt, energy, iterations) for the WHOLE batch on
    every rank (CUDA tensors on the rank's device)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    r = torch.as_tensor(r_batch, dtype=torch.float64) if not isinstance(r_batch, torch.Tensor) else r_batch
    B = int(r.shape[0])
    lo, hi = B * rank // world, B * (rank + 1) // world
    d, e, it = batched_minimize(r[lo:hi], q_core, q_shell, k, thole_u, tol=tol, maxiter=maxiter)
    dev = d.device
    per = (B + world - 1) // world
    pad = lambda t: torch.cat([t, torch.zeros((per - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)])
    outs = []
    for t in (d, e, it):
        bufs = [torch.empty((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for _ in range(world)]
        dist.all_gather(bufs, pad(t), group=group)
        outs.append(torch.cat([bufs[g][: B * (g + 1) // world - B * g // world] for g in range(world)]))
    return tuple(outs)
