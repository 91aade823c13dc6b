"""Prints the per-phase cycle counters of the tcgen05 fused kernel (DESMO_TC_DEBUG=1)."""
import ctypes, os, sys
os.environ["DESMO_TC_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from desmo_b200 import DesmoEngine, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 8
e = DesmoEngine(n, 1000, 2, 4, omega_init=10.0, device=torch.device("cuda:0"), path=2)
g = torch.Generator(device="cuda").manual_seed(0)
e.P[:, :n] = torch.randn(4, n, device="cuda", generator=g) / n ** 0.5
e.U = torch.randn(1000, e.ld, device="cuda", generator=g)
e.build_w(False)
for _ in range(3):
    e.fused_residual_grad()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); e.fused_residual_grad(); ev1.record(); torch.cuda.synchronize()
print("kernel+reduce ms", ev0.elapsed_time(ev1), "tiles/CTA", n / 128 / 148)
out = np.zeros(16384, np.uint64)
_lib.check(e.lib.desmo_debug_timers(ctypes.byref(e.shape), e.workspace.data_ptr(), out.ctypes.data_as(ctypes.c_void_p), out.size))
t = out[:148 * 32].reshape(148, 32).astype(np.float64)
nst = n / 128 / 148 * 8
names = ["mma:wait W_FULL", "mma:wait REC_EMPTY", "mma:wait GT_FULL", "mma:wait R_FULL", "mma:wait D_EMPTY", "", "", "",
         "epi:wait REC_FULL + U_FULL (probed together)", "epi:phase A (tcgen05.ld, residual, split)", "epi:wait R_EMPTY", "epi:after R_FULL (D drain / next library terms)", "epi:  R_s stores", "", "epi:next tile's TMEM library store", "epi:  fence + arrive R_FULL", "epi:total",
         "", "", "", "prodU(warp 2):wait U_EMPTY", "prodU(warp 2):total", "prodU(warp 3):wait U_EMPTY", "prodU(warp 3):total",
         "epi:wait LAT_FULL", "epi:library terms"]
for i, nm in enumerate(names):
    if nm:
        print(f"{nm:48s} mean {t[:, i].mean() / nst:9.0f} cycles/slab-tile   (min {t[:, i].min() / nst:8.0f}, max {t[:, i].max() / nst:8.0f})")

w = out[8192:8192 + 128].reshape(16, 8).astype(np.float64) / nst
print("per-warp (CTA 0)  e: q h | wait REC+U | phase A (-) | wait R_EMPTY | after | TMEM lib")
for e_ in range(16):
    print(f"  warp {e_:2d}: q{e_ & 3} h{e_ >> 2} | {w[e_,0]:7.0f} | {w[e_,1]:7.0f} ({w[e_,4]:5.0f}) | {w[e_,2]:6.0f} | {w[e_,3]:6.0f} | {w[e_,6]:5.0f}")
