"""Mixed difference-form kernel: accuracy against the fp64 path and timing, one line per call.
UPOL_DIFF_VARIANT selects the launch variant (pair_diff.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
if nmol == 20000:
    s, d = bench.make_workload()
else:
    s = synth.imidazole_box(nmol, seed=3); d = synth.random_displacements(s, seed=4)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
e64, g64 = sysm.energy_grad(D, precision="fp64")
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): r = f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r
t_eg, (em, gm) = timeit(lambda: sysm.energy_grad(D, precision="mixed"))
t_g, (_, gf) = timeit(lambda: sysm.energy_grad(D, precision="mixed", want_energy=False))
def step64():
    sysm.invalidate_static(); return sysm.energy_grad(D, precision="fp64")
t_64, _ = timeit(step64, 5)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
l2 = lambda a, b: float(torch.linalg.norm(a - b) / torch.linalg.norm(b))
inter = sysm.interactions(False); inter_s = sysm.interactions(True)
print(f"variant {os.environ.get('UPOL_DIFF_VARIANT','0')} old={'UPOL_OLD_MIXED' in os.environ} nmol {nmol}: "
      f"mixed E+g {t_eg:.3f} ms ({(20*inter+12*inter_s)/t_eg/1e9:.1f} TF alg), field-only {t_g:.3f} ms ({20*inter/t_g/1e9:.1f} TF alg), fp64 full {t_64:.3f} ms | "
      f"E rel err {abs(float(em[0])-float(e64[0]))/abs(float(e64[0])):.2e} (inter {abs(float(em[1])-float(e64[1]))/abs(float(e64[1])):.2e}) "
      f"g max {rel(gm, g64):.2e} l2 {l2(gm, g64):.2e} | field-only g max {rel(gf, g64):.2e} l2 {l2(gf, g64):.2e}", flush=True)
