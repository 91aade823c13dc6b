"""Times one fused call (build_w excluded) per path for a few (n, m, r, p) shapes: which implementation should AUTO pick?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from desmo_b200 import DesmoEngine

def run(n, m, r, p, path, iters=20):
    try:
        e = DesmoEngine(n, m, p, r, omega_init=10.0, device=torch.device("cuda:0"), path=path)
    except Exception as ex:
        return None
    g = torch.Generator(device="cuda").manual_seed(0)
    e.P[:, :n] = torch.randn(r, n, device="cuda", generator=g) / n ** 0.5
    e.U = torch.zeros(m, e.ld, device="cuda")
    e.U[:, :n] = torch.randn(m, n, device="cuda", generator=g)
    e.build_w(False)
    for _ in range(3):
        e.fused_residual_grad()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        e.fused_residual_grad()
    ev1.record(); torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / iters, e.K

shapes = [(3961, 1001, 4, 3), (3961, 1001, 8, 2), (3961, 1001, 8, 3), (16384, 1000, 4, 2), (16384, 1000, 32, 2), (27000, 1000, 4, 2),
          (1 << 20, 1000, 4, 3), (1 << 20, 1000, 8, 2), (1 << 19, 1000, 8, 3), (1 << 18, 1000, 32, 2), (1 << 20, 1000, 4, 2)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (n, m, r, p) in shapes:
    out = []
    for path in (1, 2, 3):
        res = run(n, m, r, p, path, iters=5 if n > 100000 else 20)
        out.append("   --   " if res is None else f"{res[0]:8.3f}")
        K = res[1] if res else K
    flops = 6.0 * K * n * m
    best = min(float(v) for v in out if v.strip() != "--")
    print(f"n={n:8d} m={m} r={r:2d} p={p} K={K:4d}  ms: ffma {out[0]}  fused-tc {out[1]}  gemm {out[2]}   best: {flops / best / 1e9:8.1f} TFLOP/s fp32-equivalent, {4.0 * n * m / best / 1e6:7.1f} GB/s of snapshots", flush=True)
