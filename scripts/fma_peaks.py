import sys; sys.path.insert(0, '/root/repo')
from u_pol_b200.engine import measure_fma_peak
for m in (True, False, "f32x2"):
    print(m, f"{measure_fma_peak(m)/1e12:.2f} TFLOP/s")
