"""Multi-GPU check (run under torchrun): row-sharded energy/gradient and minimiser against the
single-GPU result computed on rank 0's own device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
s = synth.imidazole_box(nmol, seed=3)
D = synth.random_displacements(s, seed=4)
full = UindSystem.from_drude_system(s, device=lr)
e1, g1 = full.energy_grad(D)
d1, i1 = full.minimize(tol=1e-4)
sh = UindSystem.from_drude_system(s, device=lr)
sh.attach_process_group(p2p=os.environ.get("UPOL_NO_P2P") is None)
e2, g2 = sh.energy_grad(D)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
d2, i2 = sh.minimize(tol=1e-4)
torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
t0 = time.perf_counter(); full.minimize(tol=1e-4); torch.cuda.synchronize(); t_full = time.perf_counter() - t0
rel_e = abs(float(e2[0]) - float(e1[0])) / abs(float(e1[0]))
rel_g = float((g2 - g1).abs().max() / g1.abs().max())
rel_m = abs(i2["U_ind"] - i1["U_ind"]) / abs(i1["U_ind"])
dd = float((d2 - d1).abs().max())
print(f"rank {rank}/{world}: dE {rel_e:.2e} dg {rel_g:.2e} dUmin {rel_m:.2e} dd {dd:.2e} iters {i1['iterations']}/{i2['iterations']} "
      f"evals {i1['evaluations']}/{i2['evaluations']} t_full {t_full:.3f}s t_sharded {t_sh:.3f}s", flush=True)
assert rel_e < 1e-11 and rel_g < 1e-12 and rel_m < 1e-10, "sharded result differs"
# energy-only and gradient-only calls, repeated (epoch parity of the peer-memory exchange), then the sharded host path
_, g1f = full.energy_grad(D, want_energy=False)      # the field-only pass (second-order cube) of the unsharded system
for rep in range(3):
    e3, _ = sh.energy_grad(D, want_grad=False)
    _, g3 = sh.energy_grad(D, want_energy=False)
    de3, dg3, dg3f = abs(float(e3[0]) - float(e1[0])) / abs(float(e1[0])), float((g3 - g1).abs().max() / g1.abs().max()), float((g3 - g1f).abs().max() / g1.abs().max())
    assert de3 <= 1e-11 and dg3 < 1e-10 and dg3f < 1e-12, (de3, dg3, dg3f)
em, gm = sh.energy_grad(D, precision="mixed")
assert abs(float(em[0]) - float(e1[0])) <= 1e-5 * abs(float(e1[0])) and float((gm - g1).abs().max() / g1.abs().max()) < 1e-5
r_host = np.ascontiguousarray(s.r_core.reshape(-1)); d_host = np.ascontiguousarray(np.asarray(D).reshape(-1))
for rep in range(3):
    eh, gh, (slo, shi) = sh.energy_grad_host_sharded(d_host, r_core=r_host if rep != 1 else None)
    gt = torch.as_tensor(gh).to(g1.device)
    dist.all_reduce(gt)                       # the caller assembles the slices (zeros elsewhere)
    assert abs(eh[0] - float(e1[0])) <= 1e-11 * abs(float(e1[0])), (eh[0], float(e1[0]))
    assert float((gt - g1).abs().max() / g1.abs().max()) < 1e-12
    assert np.all(gh[:slo] == 0) and np.all(gh[shi:] == 0)
print(f"rank {rank}: peer-memory energy/gradient exchange and sharded host path ok (sites {slo}:{shi})", flush=True)
cg1 = full.core_gradient(D); cg2 = sh.core_gradient(D)
rel_c = float((cg1 - cg2).abs().max() / cg1.abs().max())
assert rel_c < 1e-10, rel_c
print(f"rank {rank}: sharded core gradient ok ({rel_c:.1e})", flush=True)
# config 5: replicas only
from u_pol_b200.engine import batched_minimize, batched_minimize_replicas
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dimers312.npz"))
al, qs = z["alpha"], z["q_shell"]
kk = np.where(al == 0, 0.0, 138.93545764438198 * qs**2 / np.where(al == 0, 1.0, al))
dA, eA, itA = batched_minimize_replicas(z["r_core"], z["q_core"], qs, kk, z["thole_u"], tol=1e-6)
dB, eB, itB = batched_minimize(z["r_core"], z["q_core"], qs, kk, z["thole_u"], tol=1e-6)
assert torch.equal(eA[:, 0], eB[:, 0]) and torch.equal(itA, itB) and eA.shape[0] == 312
print(f"rank {rank}: batched replicas ok ({eA.shape[0]} systems)", flush=True)
dist.barrier(); dist.destroy_process_group()
