#!/usr/bin/env python
"""Benchmark of the U_ind hot path (BASELINE.json: "U_ind energy+grad pair-interactions/s and
minimise wall-time at N=100k Drudes").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference algorithm on host cores

Workload (BASELINE configs[3], SURVEY 8d "config 4"): synthetic imidazole-parameter polarisable
liquid, 20 000 molecules -> N = 180 000 sites, N_D = 100 000 Drudes, seed 4, d ~ N(0, 0.002 nm)
on Drude sites.  It fits one GPU, so it is the N=1 workload too; at N>1 the Drude rows are
sharded by molecule blocks (all columns replicated) and the total work is fixed ("strong").

A step = one full U_ind + dU/dd evaluation as the reference's jitted value_and_grad does it:
the d-independent static part (U_coul_static and the f(R) terms) is RECOMPUTED every step, the
main core/shell pass, the Thole intra pass, the spring term, the reductions, and (N>1) the
all-reduce of the energy vector and the all-gather of the gradient rows.
metric/value = reference site-pairs per second, P = N(N-1) ordered site pairs per evaluation
(the unit in which the JAX/CPU path and the CUDA path are comparable).

JSON line: see README/DESIGN.md; `roofline.bound` is "fp64" because this path is bound by the
FP64 pipe (north star), whose peak is measured on the box by an FMA-chain microbenchmark
(MEASURED_PEAKS.json has only HBM and bf16-GEMM peaks).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NMOL = int(os.environ.get("UPOL_BENCH_NMOL", "20000"))
SEED = 4
METRIC = "uind_energy_grad_site_pairs_per_s"
UNIT = "site-pairs/s"
WORKLOAD = f"imidazole_liquid_{NMOL}mol_N{9 * NMOL}_ND{5 * NMOL}_seed{SEED}"
FLOP_PER_INTERACTION = 20
FLOP_PER_STATIC_INTERACTION = 12


def make_workload():
    from u_pol_b200 import synth

    cache = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"upol_bench_{WORKLOAD}.npz")
    s = None
    if os.path.isfile(cache):
        try:
            z = np.load(cache)
            m = synth.load_monomer("imidazole")
            tile = lambda v: np.ascontiguousarray(np.broadcast_to(v, (NMOL,) + v.shape))
            from u_pol_b200.topology import DrudeSystem
            s = DrudeSystem(r_core=z["r"], q_core=tile(m["q_core"]), q_shell=tile(m["q_shell"]),
                            alpha=tile(m["alpha"]), thole_u=tile(m["thole_u"]), site_names=list(m["site_names"]))
        except Exception:
            s = None
    if s is None:
        s = synth.imidazole_box(NMOL, seed=SEED)
        try:
            tmp = cache + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, r=s.r_core)
            os.replace(tmp, cache)
        except Exception:
            pass
    d = synth.random_displacements(s, seed=SEED + 1, sigma=0.002)
    return s, d


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled during the timed region (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonGpuIdle", 0x1): "gpu_idle",
            getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSyncBoost", 0x10): "sync_boost",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_sample(s, d, target_seconds):
    """Time the C restatement of the reference on the host cores for a bounded row sample.
    Returns (site_pairs_per_s, cores, description)."""
    from oracle import c_oracle

    c_oracle.use_all_cores()
    n = s.nmol * s.natoms
    rows = 64 * s.natoms
    t0 = time.perf_counter()
    c_oracle.energy_grad(s.r_core, d, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=rows)
    t_probe = max(time.perf_counter() - t0, 1e-4)
    rows = int(min(n, max(rows, rows * target_seconds / t_probe)))
    rows -= rows % s.natoms
    t0 = time.perf_counter()
    c_oracle.energy_grad(s.r_core, d, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=rows)
    dt = time.perf_counter() - t0
    return rows * (n - 1) / dt, c_oracle.max_threads(), rows, dt


def run_reference(args):
    """Reference arm: the reference's algorithm (C restatement of polarization.py, oracle/uind_oracle.c)
    on all host threads; JAX/jaxopt are not installable in this image."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.use_all_cores()          # torchrun exports OMP_NUM_THREADS=1 to its workers
    s, d = make_workload()
    n = s.nmol * s.natoms
    # size one step to ~2 s of host work
    rows = 64 * s.natoms
    t0 = time.perf_counter()
    c_oracle.energy_grad(s.r_core, d, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=rows)
    t_probe = max(time.perf_counter() - t0, 1e-4)
    rows = int(min(n, max(rows, rows * 2.0 / t_probe)))
    rows -= rows % s.natoms
    for _ in range(args.warmup):
        c_oracle.energy_grad(s.r_core, d, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.energy_grad(s.r_core, d, s.q_core, s.q_shell, s.k, s.thole_u, row_lo=0, row_hi=rows)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows * (n - 1) / dt
    cores = c_oracle.max_threads()
    sample = f"{rows} of {n} site rows x all {n} columns per step (E + dE/dd), oracle/uind_oracle.c, OpenMP"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sites": n, "drudes": int((s.q_shell != 0).sum()),
                   "note": "reference algorithm restated in C (JAX/jaxopt/OpenMM not installable here); "
                           "each step is a bounded row sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from u_pol_b200 import _lib as L
    from u_pol_b200.engine import UindSystem, measure_fma_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    s, d_host = make_workload()
    n = s.nmol * s.natoms
    nd = int((s.q_shell != 0).sum())
    sysm = UindSystem.from_drude_system(s, device=local_rank)
    exchange = "none"
    if world > 1:
        sysm.attach_process_group()
        exchange = "peer-memory stores (CUDA IPC/NVLink) fused into the update kernel" if getattr(sysm, "p2p", False) else "nccl"
    d_dev = torch.as_tensor(d_host.reshape(-1)).to(dev)
    pairs_per_step = n * (n - 1)
    inter_local = sysm.interactions(False)
    inter_static_local = sysm.interactions(True)
    n_d_mol = int((s.q_shell[0] != 0).sum())
    inter_total = s.nmol * n_d_mol * ((n - s.natoms) + (nd - n_d_mol))          # I_inter, SURVEY 8d
    inter_static_total = s.nmol * n_d_mol * (n - s.natoms)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def step():
        sysm.invalidate_static()              # do ALL the work of an evaluation every step
        return sysm.energy_grad(d_dev, precision="fp64")

    # ---- device-resident steps -----------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        e, g = step()
    barrier()
    sysm.set_kernel_timing(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(float(i))             # L2 flush between timed iterations (not timed)
            if world > 1:
                dist.barrier()
            ev[i][0].record()
            e, g = step()
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_per_step = max_over_ranks(float(np.sum(ms_steps)) / args.steps)
    k_ms, k_n = sysm.kernel_time()
    sysm.set_kernel_timing(False)
    value = pairs_per_step / (ms_per_step * 1e-3)
    e_host = e.cpu().numpy()

    # ---- end to end through the C ABI with host buffers (H2D of r and d, D2H of E and grad) --
    # page-locked host buffers (the contract: inputs come from pinned host memory)
    r_pin = torch.from_numpy(np.ascontiguousarray(s.r_core.reshape(-1))).pin_memory()
    d_pin = torch.from_numpy(np.ascontiguousarray(d_host.reshape(-1))).pin_memory()
    e_pin = torch.empty(L.E_SIZE, dtype=torch.float64).pin_memory()
    g_pin = torch.empty(3 * n, dtype=torch.float64).pin_memory()
    r_host, d_flat, e_out, g_out = r_pin.numpy(), d_pin.numpy(), e_pin.numpy(), g_pin.numpy()
    lib = L.lib()

    def e2e_step():
        L.check(lib.upol_system_set_positions(sysm._h, r_host.ctypes.data, None))
        L.check(lib.upol_uind_energy_grad_host(sysm._h, d_flat.ctypes.data, L.FP64, e_out.ctypes.data, g_out.ctypes.data))

    with torch.cuda.device(dev):
        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / n_e2e * 1e3)
    e2e_value = pairs_per_step / (e2e_ms * 1e-3)
    assert abs(e_out[0] - e_host[0]) <= 1e-9 * abs(e_host[0]), (e_out[0], e_host[0])

    # ---- minimisation wall time (second half of the metric) --------------------------------
    minimise = {}
    for prec in ("fp64", "mixed"):
        sysm.minimize(tol=1e-4, precision=prec, maxiter=3)           # warm the workspace
        barrier()
        t0 = time.perf_counter()
        _, info = sysm.minimize(tol=1e-4, precision=prec)
        torch.cuda.synchronize()
        wall = max_over_ranks(time.perf_counter() - t0)
        minimise[prec] = {"wall_s": wall, "iterations": info["iterations"], "evaluations": info["evaluations"],
                          "U_ind": info["U_ind"], "gnorm": info["gnorm"], "flags": info["flags"], "tol": 1e-4}

    # ---- roofline of the dominant kernel (pair_fp64_kernel<POT,GRAD>) ----------------------
    peak64 = measure_fma_peak(True, local_rank)
    peak32 = measure_fma_peak(False, local_rank)
    k_ms_avg = k_ms / max(k_n, 1)
    achieved = FLOP_PER_INTERACTION * inter_local / (k_ms_avg * 1e-3) / 1e12
    traffic = pipe_ncu = None
    tfile = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tfile):
        try:
            tj = json.load(open(tfile))
            traffic = tj.get("pair_fp64_kernel_dram_bytes_per_launch")
            pipe_ncu = tj.get("pair_fp64_kernel_fp64_pipe_active_pct")      # from the committed ncu capture, not live
        except Exception:
            traffic = pipe_ncu = None
    # mixed-precision field kernel, for context
    sysm.set_kernel_timing(True)
    for _ in range(3):
        sysm.energy_grad(d_dev, precision="mixed", want_energy=False)
    torch.cuda.synchronize()
    sysm.kernel_time()
    for _ in range(5):
        sysm.energy_grad(d_dev, precision="mixed", want_energy=False)
    mx_ms, mx_n = sysm.kernel_time()
    sysm.set_kernel_timing(False)
    mixed_tflops = FLOP_PER_INTERACTION * inter_local / (mx_ms / max(mx_n, 1) * 1e-3) / 1e12

    # kernels of ours launched per step: static pass (pair, finalize, reduce), main pass (gather,
    # prepare, pair, finalize, reduce, assemble, scatter) [+ assemble after the all-reduce]
    launches_per_step = 10 + (1 if world > 1 else 0)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, rows, dt = cpu_sample(s, d_host, target_seconds=12.0)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{rows} of {n} site rows x all columns, one E+grad pass, {dt:.1f} s, "
                                  "oracle/uind_oracle.c (C restatement of polarization.py, OpenMP)"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sites": n, "drudes": nd, "parallelism": f"row-shard x{world}",
                       "l2": "flushed between timed steps (256 MiB write, untimed)",
                       "step": "full U_ind + grad incl. recomputed static part",
                       "minimiser_exchange": exchange},
            "interactions_per_s": (inter_total + inter_static_total) / (ms_per_step * 1e-3),
            "U_ind": float(e_host[0]),
            "wall_s_timed_region": t_wall,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(r_host.nbytes + d_flat.nbytes),
                    "d2h_bytes_per_step": int(g_out.nbytes + e_out.nbytes)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clk.summary(),
            "roofline": {"bound": "fp64", "kernel": "pair_fp64_kernel<POT,GRAD>", "achieved": achieved,
                         "peak": peak64 / 1e12, "unit": "TFLOP/s", "frac": achieved / (peak64 / 1e12),
                         "peak_source": "measured on this box: upol_measure_fma_peak (DFMA chain); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "kernel_ms": k_ms_avg, "launches_timed": k_n,
                         "interactions_per_launch": inter_local, "flop_per_interaction": FLOP_PER_INTERACTION,
                         # the kernel issues 18.3 FP64-pipe instructions per interaction (SASS count: DFMA, DADD, DMUL;
                         # DESIGN section 5): share of the DFMA-chain issue rate those instructions take up
                         "fp64_instr_per_interaction": 18.3,
                         "fp64_issue_frac": 18.3 * inter_local / (k_ms_avg * 1e-3) / (peak64 / 2.0),
                         "fp64_pipe_active_pct_ncu": pipe_ncu,
                         "traffic": traffic,
                         "mixed_field_kernel": {"achieved": mixed_tflops, "peak": peak32 / 1e12,
                                                "frac": mixed_tflops / (peak32 / 1e12), "unit": "TFLOP/s"}},
            "minimise": minimise,
        }
        if cpu_baseline is not None:
            out["cpu_baseline"] = cpu_baseline
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
