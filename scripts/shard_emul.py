"""One rank's share of an 8-way row-sharded evaluation on ONE GPU (upol_system_set_row_range, no exchange): where do
the microseconds between the pair kernel and the step go?  Prints the step time (CUDA events, L2 flushed between steps)
next to the pair-kernel time; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from u_pol_b200.engine import UindSystem
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
s, d = bench.make_workload()
sysm = UindSystem.from_drude_system(s)
sysm.set_row_range(0, s.nmol // parts)
D = torch.as_tensor(d.reshape(-1)).cuda()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
def step():
    sysm.invalidate_static()
    return sysm.energy_grad(D)
for _ in range(3): step()
torch.cuda.synchronize()
ts = []
for i in range(10):
    flush.fill_(float(i))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); step(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
sysm.set_kernel_timing(True)
for _ in range(5): step()
ms, n = sysm.kernel_time()
print(f"1/{parts} of the rows: step {sum(ts)/len(ts)*1e3:.1f} us (min {min(ts)*1e3:.1f}), pair kernel {ms/n*1e3:.1f} us")
