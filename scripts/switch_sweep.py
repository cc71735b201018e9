"""Mixed-precision minimiser: where to hand over from the fp32 field to the fp64 field (UPOL_MIXED_SWITCH, relative to
||g(d=0)||).  Run once per value (the library reads the variable once): prints wall time, iterations, minimum."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from u_pol_b200.engine import UindSystem
for cfg in (3, 4):
    s, d, name = bench._config_workload(cfg)
    sysm = UindSystem.from_drude_system(s)
    out = []
    for prec in ("fp64", "mixed"):
        sysm.minimize(tol=1e-4, precision=prec, maxiter=3)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _, info = sysm.minimize(tol=1e-4, precision=prec)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        out.append(f"{prec} {t*1e3:8.2f} ms it {info['iterations']} ev {info['evaluations']} U {info['U_ind']:.9f} g {info['gnorm']:.2e} fl {info['flags']}")
    print(f"switch {os.environ.get('UPOL_MIXED_SWITCH', '1e-4 (default)')} config {cfg}: " + " | ".join(out), flush=True)
