"""Times the fused tcgen05 kernel of whichever library DESMO_B200_LIB points at (kernel experiments; results may be wrong on purpose)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desmo_b200 import DesmoEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 64
e = DesmoEngine(n, 1000, 2, 4, omega_init=10.0, device=torch.device("cuda:0"), path=2)
g = torch.Generator(device="cuda").manual_seed(0)
e.P[:, :n] = torch.randn(4, n, device="cuda", generator=g) / n ** 0.5
e.U = torch.randn(1000, e.ld, device="cuda", generator=g)
e.build_w(False)
for _ in range(3):
    e.fused_residual_grad()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(10):
    e.fused_residual_grad()
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 10
st = n / 128 / 148 * 8
print(f"{os.environ.get('DESMO_B200_LIB', 'default'):45s} fused call {ms:7.3f} ms  = {ms * 1e-3 * 1.965e9 / st:7.0f} cycles/slab-tile @1965MHz")
