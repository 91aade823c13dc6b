"""Timeline of CTA 0 of the tcgen05 fused kernel over a window of slab-tiles (DESMO_TC_DEBUG=1): MMA issuer, four epilogue warps, one
U producer.  Prints the merged event list (cycles relative to the first event) and per-slab-tile intervals."""
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
out = np.zeros(16384, np.uint64)
_lib.check(e.lib.desmo_debug_timers(ctypes.byref(e.shape), e.workspace.data_ptr(), out.ctypes.data_as(ctypes.c_void_p), out.size))
TAGS = {1: "g1:begin", 2: "g1:W_FULL ok", 3: "g1:REC_EMPTY ok", 4: "g1:GT_FULL ok", 5: "g1:issued", 6: "g34:begin", 7: "g34:R_FULL ok",
        9: "g3 issued", 10: "g4 q0 issued", 11: "g4 q1 issued", 12: "g4 q2 issued", 13: "g4 q3 issued",
        20: "wait REC_FULL", 21: "REC_FULL+U_FULL ok", 22: "tmem ld issued", 23: "tmem ld done", 24: "residual done/REC_EMPTY", 25: "split done",
        26: "R_EMPTYQ ok", 27: "R_s stored", 28: "fence+R_FULL arrive", 29: "after done", 30: "next library -> TMEM",
        50: "REC_FULL completes (G1 done)", 51: "R_EMPTYQ3 completes (G4 done)"}
ROLE = ["mma", "epi q0h0", "epi q1h1", "epi q2h2", "epi q3h3", "pipe"]
ev = []
for log in range(6):
    for w in out[8320 + log * 384: 8320 + (log + 1) * 384]:
        w = int(w)
        if w:
            ev.append((w & 0xffffffffff, log, (w >> 56) & 0xff, (w >> 40) & 0xffff))
ev.sort()
t0 = ev[0][0]
for t, log, tag, it in ev:
    print(f"{t - t0:8d}  it {it:3d}  {ROLE[log]:9s} {'    ' * log}{TAGS.get(tag, tag)}")
