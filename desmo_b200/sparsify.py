"""Post-hoc sparsification (CYL:624-692,1184-1270; TURB:1166-1245): term norms -> threshold sweep -> active mask."""
from __future__ import annotations

from typing import List, Tuple

import torch


def threshold_sweep(model, snapshot_norm2: float, thresholds=None) -> List[Tuple[float, float, int, torch.Tensor]]:
    """For every threshold: zero the gates whose term norm is below it (CYL:1228-1238), evaluate
    ||X - recon^T|| / ||X|| (CYL:1257) and count the surviving non-zero gates (CYL:1260-1265).
    Returns [(threshold, relative_error, n_active, mask)], restoring the gates afterwards."""
    e = model.engine
    if thresholds is None:
        thresholds = [10.0 ** (-4 + 0.5 * i) for i in range(14)]  # 10^-4 .. 10^2.5, half-decade steps (CYL:1213)
    norms = e.term_norms()
    saved = e.gates.clone()
    out = []
    for thr in thresholds:
        e.gates.copy_(saved)
        mask = (norms >= thr) & (saved != 0)
        e.gates.mul_(mask.to(e.gates.dtype))
        err = (e.residual_norm2() / snapshot_norm2) ** 0.5
        out.append((float(thr), float(err), int(mask.sum().item()), mask.clone()))
    e.gates.copy_(saved)
    return out


def greedy_removal(model, snapshot_norm2: float) -> List[Tuple[int, float]]:
    """TURB:1166-1245: sort terms by norm, remove the smallest one at a time, re-evaluate the relative error."""
    e = model.engine
    norms = e.term_norms()
    order = torch.argsort(norms)
    saved = e.gates.clone()
    out = []
    for k in order.tolist():
        e.gates[k] = 0.0
        out.append((k, float((e.residual_norm2() / snapshot_norm2) ** 0.5)))
    e.gates.copy_(saved)
    return out
