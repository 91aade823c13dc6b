"""Post-hoc sparsification (CYL:624-692,1184-1270; TURB:1166-1245): term norms -> threshold sweep -> active mask."""
from __future__ import annotations

from typing import List, Tuple

import torch


def threshold_sweep(model, snapshot_norm2: float, thresholds=None) -> List[Tuple[float, float, int, torch.Tensor]]:
    """For every threshold: zero the gates whose term norm is below it (CYL:1228-1238), evaluate
    ||X - recon^T|| / ||X|| (CYL:1257) and count the surviving non-zero gates (CYL:1260-1265).
    Returns [(threshold, relative_error, n_active, mask)], restoring the gates afterwards."""
    e = model.engine
    if hasattr(model, "sync_parameters"):
        model.sync_parameters()
    if thresholds is None:
        thresholds = [10.0 ** (-4 + 0.5 * i) for i in range(14)]  # 10^-4 .. 10^2.5, half-decade steps (CYL:1213)
    norms = e.term_norms()  # the reference's semantics (raw phi_list; Fourier column quirk), see DesmoEngine.term_norms
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


def removal_order(norms: torch.Tensor, T: int, r: int) -> List[int]:
    """Packed K indices in the order the reference's sweep removes them: the list is built as the polynomial terms, then
    (sin_i, cos_i, tanh_i) per mode (TURB:1173-1181), and sorted by norm with a stable sort (TURB:1183)."""
    vals = norms.tolist()
    listed = list(range(T)) + [T + b * r + i for i in range(r) for b in range(3)]
    return sorted(listed, key=lambda k: vals[k])


def greedy_removal(model, snapshot_norm2: float) -> List[Tuple[int, float, int]]:
    """TURB:1166-1245: for step = 0..K zero the gates of the ``step`` smallest-norm terms, evaluate ||X - recon^T|| / ||X||
    (TURB:1227) and count the non-zero gates left (TURB:1229-1234).  One gradient-free fused pass per step on the resident
    snapshot instead of a DataLoader round trip + host numpy.  Returns [(step, relative_error, nonzero_terms)], gates restored."""
    e = model.engine
    if hasattr(model, "sync_parameters"):
        model.sync_parameters()
    order = removal_order(e.term_norms(), e.T, e.r)
    saved = e.gates.clone()
    out = []
    for step in range(e.K + 1):
        if step:
            e.gates[order[step - 1]] = 0.0
        out.append((step, float((e.residual_norm2() / snapshot_norm2) ** 0.5), int((e.gates != 0).sum().item())))
    e.gates.copy_(saved)
    return out
