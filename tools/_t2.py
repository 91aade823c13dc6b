import sys, torch, numpy as np
sys.path.insert(0, ".")
from desmo_b200 import DesmoEngine
n, m = int(sys.argv[1]), 1000
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
g = torch.Generator(device="cuda").manual_seed(0)
P = torch.randn(4, n, device="cuda", generator=g) / n ** 0.5
rows = torch.randn(27, m, device="cuda", generator=g)
U = torch.randn(m, n, device="cuda", generator=g)
out = {}
for path in (1, 2):
    e = DesmoEngine(n, m, 2, 4, omega_init=10.0, device=torch.device("cuda:0"), path=path)
    e.P[:, :n] = P; e.rows[:, :m] = rows
    e.U = torch.zeros(m, e.ld, device="cuda"); e.U[:, :n] = U
    e.build_w(False); e.fused_residual_grad(); torch.cuda.synchronize()
    out[path] = (e.red.clone(), e.dphi[:, :n].clone(), e.Kp * e.mld)
    del e
r1, d1, ec = out[1]; r2, d2, _ = out[2]
print(f"n={n}: E rel diff {rel(r2[:ec], r1[:ec]):.3e}  loss rel diff {abs(float(r2[ec]) - float(r1[ec])) / float(r1[ec]):.3e}  dphi rel diff {rel(d2, d1):.3e}  gram {rel(r2[ec+1:ec+17], r1[ec+1:ec+17]):.2e} domega {rel(r2[ec+17:], r1[ec+17:]):.2e}")
# signed bias of E: mean of (tc - fp32) * sign(fp32) relative to mean |fp32|
E1, E2 = r1[:ec].double(), r2[:ec].double()
print("  E signed relative bias:", float(((E2 - E1) * torch.sign(E1)).sum() / E1.abs().sum()))
