"""Minimiser polling interval (check_every) on the small configs: wall time per minimisation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from u_pol_b200.engine import UindSystem
for cfg in (2, 3):
    s, d, name = bench._config_workload(cfg)
    sysm = UindSystem.from_drude_system(s)
    for prec in ("fp64", "mixed"):
        out = []
        for ce in (1, 2, 3, 4, 6, 8):
            for _ in range(3): sysm.minimize(tol=1e-4, precision=prec, check_every=ce)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10): _, info = sysm.minimize(tol=1e-4, precision=prec, check_every=ce)
            torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
            out.append(f"{ce}: {t*1e3:.3f} ms")
        print(f"config {cfg} {prec} ({info['evaluations']} evals): " + "  ".join(out), flush=True)
