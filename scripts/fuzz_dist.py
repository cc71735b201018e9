"""Multi-GPU fuzz (run under torchrun): random boxes and random topologies, row-sharded with and
without the peer-memory exchange, against the same system evaluated unsharded on the rank's own GPU.
Also exercises create / attach / destroy many times in one process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from test_gpu_random_topologies import _random_system
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(77)           # same stream on every rank
bad = 0
worst = dict(dE=0.0, dg=0.0, dmin=0.0)
for case in range(n_cases):
    kind = case % 3
    if kind == 2:
        seed = int(rng.integers(0, 10**6))
        r, qc, qs, k, th, D = _random_system(seed)
        mk = lambda: UindSystem.from_compact(r, qc, qs, k, th, device=lr)
        label = f"topology seed {seed} nmol {r.shape[0]}"
    else:
        maker = synth.water_box if kind == 0 else synth.imidazole_box
        nmol = int(rng.integers(1, 700)) if case % 5 else int(rng.integers(1, 2 * world))
        seed = int(rng.integers(1, 10**6))
        s = maker(nmol, seed=seed)
        D = synth.random_displacements(s, seed=seed + 1)
        mk = lambda: UindSystem.from_drude_system(s, device=lr)
        label = f"{maker.__name__} nmol {nmol} seed {seed}"
    p2p = bool(case % 2)
    full, sh = mk(), mk()
    sh.attach_process_group(p2p=p2p)
    e1, g1 = full.energy_grad(D); e2, g2 = sh.energy_grad(D)
    scale = abs(float(e1[0])) + 1e-12 * abs(float(e1[5])) + 1e-300
    dE = abs(float(e2[0]) - float(e1[0])) / scale
    gmax = float(g1.abs().max())
    dg = float((g2 - g1).abs().max()) / gmax if gmax > 0 else float((g2 - g1).abs().max())
    # sharded host path (only this rank's slices cross PCIe) and a mixed call through the same exchange
    nsite = g1.shape[0]
    r_host = np.ascontiguousarray(np.asarray(full._keep[0]).reshape(-1)); d_host = np.ascontiguousarray(np.asarray(D, dtype=np.float64).reshape(-1))
    eh, gh, (slo, shi) = sh.energy_grad_host_sharded(d_host, r_core=r_host if case % 4 < 2 else None)
    gt = torch.as_tensor(gh).to(g1.device); dist.all_reduce(gt)
    dE = max(dE, abs(eh[0] - float(e1[0])) / scale)
    dg = max(dg, float((gt - g1).abs().max()) / (gmax if gmax > 0 else 1.0))
    em, gm = sh.energy_grad(D, precision="mixed")
    if abs(float(em[0]) - float(e1[0])) > 1e-5 * scale + 1e-9 * abs(float(e1[5])) or (gmax > 0 and float((gm - g1).abs().max()) > 1e-5 * gmax):
        dE = max(dE, 1.0)      # counted as a failure below
    _, i1 = full.minimize(tol=1e-5); _, i2 = sh.minimize(tol=1e-5)
    dmin = 0.0
    if i1["flags"] == 0 and i2["flags"] == 0 and i1["U_ind"] != 0.0:
        dmin = abs(i2["U_ind"] - i1["U_ind"]) / (abs(i1["U_ind"]) + 1e-12 * abs(float(e1[5])))
    flags_differ = (i1["flags"] == 0) != (i2["flags"] == 0)
    for kk_, v in (("dE", dE), ("dg", dg), ("dmin", dmin)):
        worst[kk_] = max(worst[kk_], v)
    if dE > 1e-10 or dg > 1e-10 or dmin > 1e-8 or flags_differ or (i2["flags"] & 8):
        bad += 1
        print(f"rank {rank} case {case} ({label}, p2p {p2p}): dE {dE:.1e} dg {dg:.1e} dmin {dmin:.1e} flags {i1['flags']}/{i2['flags']}", flush=True)
    sh.close(); full.close()
t = torch.tensor([bad], device=f"cuda:{lr}"); dist.all_reduce(t)
if rank == 0:
    print(f"{world} GPUs, {n_cases} cases: {int(t)} failing rank-cases; worst dE {worst['dE']:.2e} dg {worst['dg']:.2e} dmin {worst['dmin']:.2e}", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(min(int(t), 100))
