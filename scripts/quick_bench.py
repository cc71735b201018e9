"""Scratch timing of the pair kernels (not the contract bench; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem, measure_fma_peak

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
t = time.time(); s = synth.imidazole_box(nmol, seed=3); print("gen %.1fs" % (time.time() - t), flush=True)
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(synth.random_displacements(s, seed=1)).cuda()
print("fp64 FMA peak %.2f TFLOP/s, fp32 %.2f TFLOP/s" % (measure_fma_peak(True) / 1e12, measure_fma_peak(False) / 1e12))
inter = sysm.interactions()
gref = None
for prec, we in (("fp64", True), ("fp64", False), ("mixed", False), ("mixed", True)):
    for _ in range(3): sysm.energy_grad(D, precision=prec, want_energy=we)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): e, g = sysm.energy_grad(D, precision=prec, want_energy=we)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if gref is None: gref = g.clone()
    err = float((g - gref).abs().max() / gref.abs().max()); err2 = float(torch.linalg.norm(g - gref) / torch.linalg.norm(gref))
    print(f"{prec} energy={we}: {ms:.3f} ms/eval  {inter/ms/1e6:.1f} Ginteractions/s  alg {20*inter/ms/1e9:.2f} TFLOP/s  grad err max {err:.2e} l2 {err2:.2e}")
t = time.time(); d, info = sysm.minimize(tol=1e-4); torch.cuda.synchronize(); print("minimise fp64 %.3fs" % (time.time() - t), {k: v for k, v in info.items() if k != "energy"})
t = time.time(); d, info = sysm.minimize(tol=1e-4, precision="mixed"); torch.cuda.synchronize(); print("minimise mixed %.3fs" % (time.time() - t), {k: v for k, v in info.items() if k != "energy"})
