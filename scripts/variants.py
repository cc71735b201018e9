import os, sys, subprocess, json
code = r'''
import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
inter = sysm.interactions()
for mode in (("fp64", True), ("fp64", False)):
    for _ in range(3): sysm.energy_grad(D, precision=mode[0], want_energy=mode[1])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8): e, g = sysm.energy_grad(D, precision=mode[0], want_energy=mode[1])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 8
    print("VAR", os.environ.get("UPOL_FP64_VARIANT"), "TARGET", os.environ.get("UPOL_SPLIT_TARGET"), mode, "%.3f ms  %.1f Gint/s" % (ms, inter / ms / 1e6), flush=True)
'''
for var in ("0", "1", "2", "3", "4", "5"):
    for tgt in ("1184", "5920"):
        env = dict(os.environ, UPOL_FP64_VARIANT=var, UPOL_SPLIT_TARGET=tgt)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("\n".join(l for l in (r.stdout + r.stderr).splitlines() if l.startswith("VAR") or "rror" in l), flush=True)
