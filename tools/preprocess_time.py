"""Times desmo_preprocess on 2^20 points x 500 snapshots of 3-component fp32 input (run with PYTHONPATH=.)."""
import torch, time
from desmo_b200.engine import DesmoEngine
n, m, d = 1 << 20, 500, 3
e = DesmoEngine(n, m, 4, 2, device="cuda:0")
raw = torch.randn(m, n * d, device="cuda:0")
for _ in range(2):
    e.preprocess_snapshot(raw, d_in=3)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(5):
    e.preprocess_snapshot(raw, d_in=3)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 5
byt = n * m * (2 * d * 4 + 4)
print(f"preprocess {n}x{m} d=3 fp32: {ms:.3f} ms, {byt / ms / 1e6:.0f} GB/s algorithmic")
