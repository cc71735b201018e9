"""Periodic (Ewald) path row-sharded over GPUs (run under torchrun) against the unsharded result."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
bad = 0
for name, maker, nx, density, p2p in (("water", synth.water_box, 6, 33.4, True), ("imidazole", synth.imidazole_box, 8, 7.0, False),
                                      ("imidazole", synth.imidazole_box, 16, 7.0, True)):
    nmol = nx ** 3
    s = maker(nmol, seed=2)
    box = np.array([nx * (1.0 / density) ** (1.0 / 3.0) + 0.08] * 3)
    D = synth.random_displacements(s, seed=4)
    full = UindSystem.from_drude_system(s, device=lr); full.set_periodic(box)
    sh = UindSystem.from_drude_system(s, device=lr)
    if nx == 8:
        sh.set_periodic(box); sh.attach_process_group(p2p=p2p)       # either order of the two calls
    else:
        sh.attach_process_group(p2p=p2p); sh.set_periodic(box)
    e1, g1 = full.energy_grad(D); e2, g2 = sh.energy_grad(D)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    d2, i2 = sh.minimize(tol=1e-5); torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
    t0 = time.perf_counter(); d1, i1 = full.minimize(tol=1e-5); torch.cuda.synchronize(); t_full = time.perf_counter() - t0
    dE = abs(float(e2[0]) - float(e1[0])) / abs(float(e1[0]))
    dg = float((g2 - g1).abs().max() / g1.abs().max())
    dm = abs(i2["U_ind"] - i1["U_ind"]) / abs(i1["U_ind"])
    c1 = full.core_gradient(D); c2 = sh.core_gradient(D)
    dc = float((c2 - c1).abs().max() / c1.abs().max())
    ok = dE < 1e-11 and dg < 1e-12 and dm < 1e-10 and dc < 1e-11 and i1["flags"] == 0 and i2["flags"] == 0
    bad += not ok
    print(f"rank {rank}/{world} {name} {nmol}: dE {dE:.1e} dg {dg:.1e} dUmin {dm:.1e} dcore {dc:.1e} iterations {i1['iterations']}/{i2['iterations']} "
          f"flags {i1['flags']}/{i2['flags']} minimise {t_full*1e3:.1f} ms on 1 GPU, {t_sh*1e3:.1f} ms sharded {'OK' if ok else 'FAIL'}", flush=True)
    sh.close(); full.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(bad)
