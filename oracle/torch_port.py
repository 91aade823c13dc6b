"""CPU PyTorch port of the reference's training step -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Used only by bench.py (``cpu_baseline`` and ``--impl reference``) and tests.  The reference itself is a set of Python scripts
that cannot travel to the GPU box, so this port re-states its per-step op sequence with the same ATen operators the
reference dispatches (column-by-column ``torch.cat`` library CYL:376-434, 3r rank-1 ``mm`` outer products CYL:565-569, full
``mm`` CYL:572, ``nn.MSELoss`` CYL:722, autograd backward CYL:766, ``torch.optim.Adamax`` with the four/five param groups
CYL:592-612) so that its cost structure -- and therefore its timing -- is the reference's.  Pinned against
oracle/desmo_oracle.py (which is pinned against the reference's golden vectors) in tests/test_torch_port.py.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import desmo_oracle as orc


def library_by_cat(lat: torch.Tensor, polyorder: int) -> torch.Tensor:
    """(n, r) -> (n, T); one ``cat`` per appended column like POOL_DATA (CYL:384-431)."""
    n, r = lat.shape
    lib = torch.ones((n, 1), dtype=lat.dtype)
    for idx in orc.monomial_table(r, polyorder)[1:]:
        col = lat[:, idx[0]]
        for v in idx[1:]:
            col = col * lat[:, v]
        lib = torch.cat((lib, col.reshape(n, 1)), dim=1)
    return lib


class TorchPort(nn.Module):
    def __init__(self, prm: orc.DesmoParams, pod_modes: np.ndarray):
        super().__init__()
        self.r, self.p, self.T, self.m, self.nF = prm.r, prm.polyorder, prm.T, prm.m, prm.nF
        self.pod = pod_modes
        t = lambda a: nn.Parameter(torch.from_numpy(np.array(a, dtype=np.float32)))  # noqa: E731
        rows = prm.coefs if prm.fourier else prm.zall
        T, r = self.T, self.r
        self.phi = nn.ParameterList([t(prm.phi[i]) for i in range(r)])
        self.c = t(prm.gates[:T])
        self.z = nn.ParameterList([t(rows[j]) for j in range(T)])
        self.zs = nn.ParameterList([t(rows[T + i]) for i in range(r)])
        self.zc = nn.ParameterList([t(rows[T + r + i]) for i in range(r)])
        self.zt = nn.ParameterList([t(rows[T + 2 * r + i]) for i in range(r)])
        self.gs = nn.ParameterList([t(prm.gates[T + i]) for i in range(r)])
        self.gc = nn.ParameterList([t(prm.gates[T + r + i]) for i in range(r)])
        self.gt = nn.ParameterList([t(prm.gates[T + 2 * r + i]) for i in range(r)])
        self.om = nn.ParameterList([t(prm.omega[i]) for i in range(3 * r)])
        if prm.fourier:
            self.per = nn.ParameterList([t(prm.periods[j:j + 1]) for j in range(T + 3 * r)])
            self.tp = torch.linspace(0, prm.m, prm.m)

    def series(self, k: int, coeffs: torch.Tensor) -> torch.Tensor:
        if not self.nF:
            return coeffs
        out = coeffs[0] * torch.ones_like(self.tp)
        for h in range(1, self.nF + 1):
            out = out + (coeffs[2 * h - 1] * torch.cos(2 * torch.pi * h * self.tp / self.per[k]) +
                         coeffs[2 * h] * torch.sin(2 * torch.pi * h * self.tp / self.per[k]))
        return out

    def forward(self):
        T, r = self.T, self.r
        modes = [p * torch.from_numpy(self.pod[:, i]).type(torch.FloatTensor) for i, p in enumerate(self.phi)]  # per-call cast, CYL:538-541
        lat = torch.stack(modes, dim=1)
        theta = self.c * library_by_cat(lat, self.p)
        zvals = torch.stack([self.series(j, z) for j, z in enumerate(self.z)], dim=0)
        extra = 0
        for i in range(r):
            s = self.gs[i] * self.series(T + i, self.zs[i]).view(-1, 1) @ torch.sin(self.om[3 * i] * modes[i]).view(1, -1)
            c = self.gc[i] * self.series(T + r + i, self.zc[i]).view(-1, 1) @ torch.cos(self.om[3 * i + 1] * modes[i]).view(1, -1)
            h = self.gt[i] * self.series(T + 2 * r + i, self.zt[i]).view(-1, 1) @ torch.tanh(self.om[3 * i + 2] * modes[i]).view(1, -1)
            extra = extra + s + c + h
        recon = theta @ zvals + extra.T
        return recon.T, lat

    def optimizer(self, lrs=orc.REFERENCE_LRS):
        groups = [{"params": [self.c] + list(self.gs) + list(self.gc) + list(self.gt), "lr": lrs[0]},
                  {"params": list(self.phi), "lr": lrs[1]},
                  {"params": list(self.z) + list(self.zs) + list(self.zc) + list(self.zt), "lr": lrs[2]},
                  {"params": list(self.om), "lr": lrs[3]}]
        if self.nF:
            groups.append({"params": list(self.per), "lr": lrs[4]})
        return torch.optim.Adamax(groups, weight_decay=0.0)

    def losses(self, snapshot: torch.Tensor, beta: float, lam: float):
        recon, lat = self.forward()
        ortho = 0
        for i in range(self.r):
            for j in range(i + 1, self.r):
                ortho = ortho + torch.norm(lat[:, i] @ lat[:, j], p="fro")
        mse = nn.functional.mse_loss(recon, snapshot)
        l1 = torch.norm(self.c, p=1)
        for lst in (self.gs, self.gc, self.gt):
            for g in lst:
                l1 = l1 + torch.norm(g, p=1)
        return mse, ortho, l1, mse + beta * ortho + lam * l1

    def packed_grads(self):
        g = lambda lst: np.stack([p.grad.numpy() for p in lst])  # noqa: E731
        rows = np.concatenate([g(self.z), g(self.zs), g(self.zc), g(self.zt)])
        gates = np.concatenate([self.c.grad.numpy()] + [g(l).reshape(-1) for l in (self.gs, self.gc, self.gt)])
        out = {"phi": g(self.phi), "gates": gates, "omega": g(self.om).reshape(-1), ("coefs" if self.nF else "zall"): rows}
        if self.nF:
            out["periods"] = g(self.per).reshape(-1)
        return out


def time_steps(prm: orc.DesmoParams, pod_modes: np.ndarray, snapshot64: np.ndarray, steps: int, warmup: int, beta=1e-3, lam=1e-4,
               recollate: bool = True):
    """Seconds per step of the reference loop body on the host cores.  ``recollate`` repeats the reference's per-epoch
    fp64 -> fp32 conversion of the whole batch (DataLoader + ``.type(FloatTensor)``, CYL:707-708)."""
    import time

    model = TorchPort(prm, pod_modes)
    opt = model.optimizer()
    snap64 = torch.from_numpy(snapshot64)
    snap32 = snap64.type(torch.FloatTensor)
    loss_val = None
    for it in range(warmup + steps):
        if it == warmup:
            t0 = time.perf_counter()
        snap = snap64.clone().type(torch.FloatTensor) if recollate else snap32
        mse, ortho, l1, total = model.losses(snap, beta, lam)
        opt.zero_grad()
        total.backward()
        opt.step()
        loss_val = mse.item()  # CYL:769
    return (time.perf_counter() - t0) / steps, loss_val
