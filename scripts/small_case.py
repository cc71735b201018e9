"""Config 2 (3000 waters) and config 3 (4000 imidazoles): where does the time of one evaluation go?
Per kind of evaluation: wall time per call when every call is synchronised (latency, includes the host side),
per call in a batch of 50 enqueued back to back (device-side pipeline), and the CUDA-event time of the pair
kernel alone (the library's own events)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from u_pol_b200 import synth
from u_pol_b200.engine import UindSystem

def bench(sysm, D, label, **kw):
    inval = kw.pop("invalidate", False)
    def fn():
        if inval: sysm.invalidate_static()
        return sysm.energy_grad(D, **kw)
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn(); torch.cuda.synchronize()
    lat = (time.perf_counter() - t0) / 50
    t0 = time.perf_counter()
    for _ in range(50): fn()
    torch.cuda.synchronize()
    thr = (time.perf_counter() - t0) / 50
    sysm.set_kernel_timing(True)
    for _ in range(20): fn()
    ms, n = sysm.kernel_time()
    sysm.set_kernel_timing(False)
    print(f"  {label:34s} latency {lat*1e6:7.1f} us   back-to-back {thr*1e6:7.1f} us   pair kernel {ms/max(n,1)*1e3:7.1f} us ({n} launches)")

for name, maker, nmol, seed in (("config 2: 3000 waters", synth.water_box, 3000, 2), ("config 3: 4000 imidazoles", synth.imidazole_box, 4000, 3)):
    s = maker(nmol, seed=seed)
    sysm = UindSystem.from_drude_system(s)
    D = torch.as_tensor(synth.random_displacements(s, seed=1).reshape(-1)).cuda()
    print(name, f"N={s.nmol*s.natoms} N_D={sysm.ndrude}")
    bench(sysm, D, "fp64 E+grad, static recomputed", invalidate=True)
    bench(sysm, D, "fp64 E+grad, static cached")
    bench(sysm, D, "fp64 field only", want_energy=False)
    bench(sysm, D, "mixed E+grad", precision="mixed")
    bench(sysm, D, "mixed field only", precision="mixed", want_energy=False)
    for prec in ("fp64", "mixed"):
        sysm.minimize(tol=1e-4, precision=prec, maxiter=2)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d, info = sysm.minimize(tol=1e-4, precision=prec); torch.cuda.synchronize(); t = time.perf_counter() - t0
        print(f"  minimise {prec:5s}: {t*1e3:.2f} ms, {info['iterations']} it / {info['evaluations']} evals -> {t/info['evaluations']*1e6:.0f} us per evaluation slot, U = {info['U_ind']:.6f}")
