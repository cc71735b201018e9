import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from u_pol_b200.engine import UindSystem
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
D = torch.as_tensor(d.reshape(-1)).cuda()
_, gref = sysm.energy_grad(D, want_energy=False)
for _ in range(3): sysm.energy_grad(D, precision="mixed", want_energy=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): _, g = sysm.energy_grad(D, precision="mixed", want_energy=False)
e1.record(); torch.cuda.synchronize()
print("variant", os.environ.get("UPOL_MIXED_VARIANT", "0"), f"{e0.elapsed_time(e1)/10:.3f} ms/eval err {float((g-gref).abs().max()/gref.abs().max()):.2e}")
