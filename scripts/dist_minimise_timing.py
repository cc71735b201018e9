"""Under torchrun: minimise wall time of a row-sharded system with the peer-memory exchange vs NCCL."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for nmol in (4000, 20000):
    s = synth.imidazole_box(nmol, seed=3 if nmol == 4000 else 4)
    res = {}
    for mode in ("p2p", "nccl"):
        sh = UindSystem.from_drude_system(s, device=lr)
        sh.attach_process_group(p2p=(mode == "p2p"))
        sh.minimize(tol=1e-4, maxiter=3)
        ts = []
        for _ in range(5):
            torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
            d, info = sh.minimize(tol=1e-4); torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t[0]))
        res[mode] = (min(ts), info["iterations"], info["evaluations"], info["U_ind"], info["flags"])
        sh.close()
    if rank == 0:
        print(f"N_D={5*nmol} x{world}: p2p {res['p2p'][0]*1e3:.2f} ms, nccl {res['nccl'][0]*1e3:.2f} ms; it/evals {res['p2p'][1]}/{res['p2p'][2]}; "
              f"U {res['p2p'][3]:.9f} vs {res['nccl'][3]:.9f}; flags {res['p2p'][4]} {res['nccl'][4]}", flush=True)
dist.barrier(); dist.destroy_process_group()
