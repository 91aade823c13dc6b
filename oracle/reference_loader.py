"""Executes the reference's OWN hot-path definitions (read-only, in this container only).

TEST INFRASTRUCTURE.  The reference scripts cannot be imported verbatim (top-level ``import vtk`` / data paths,
SURVEY.md section 8c), but their hot-path definitions are pure torch.  We ``ast.parse`` the script, pick the named
``FunctionDef`` / ``ClassDef`` nodes and ``exec`` them into a namespace that supplies the module globals they read
(``device``, ``POD_modes``, ``r_DESMO``, ``polyorder``, ``t_points``, ``period_init``).  Nothing is copied into this
repo: the source is read from ``/root/reference`` at run time, so this module only works where that tree exists
(``oracle/make_golden.py`` and the pinning tests; never on the GPU box).
"""
from __future__ import annotations

import ast
import math
import os
from typing import Dict, Iterable

REF_ROOT = os.environ.get("DESMO_REFERENCE_ROOT", "/root/reference")
CYL = "DESMO/cylinder_flow/DESMO-Cylinder.py"
FCYL = "DESMO_Fourier/cylinder_flow/DESMO-Cylinder.py"

DESMO_DEFS = ("POOL_DATA", "binomial_coefficient", "calculate_number_of_terms", "DESMO", "poly_norm", "nonlinear_norm")
FOURIER_DEFS = ("POOL_DATA", "binomial_coefficient", "calculate_number_of_terms", "fourier_series", "DESMOFourier",
                "poly_norm", "nonlinear_norm")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, CYL))


def load_definitions(rel_path: str, names: Iterable[str], injected: Dict[str, object]) -> Dict[str, object]:
    import numpy as np
    import torch
    from torch import nn

    with open(os.path.join(REF_ROOT, rel_path), "r") as fh:
        tree = ast.parse(fh.read())
    wanted = set(names)
    body = [node for node in tree.body if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in wanted]
    missing = wanted - {n.name for n in body}
    if missing:
        raise RuntimeError(f"reference definitions not found in {rel_path}: {sorted(missing)}")
    ns: Dict[str, object] = {"torch": torch, "nn": nn, "np": np, "math": math, "device": torch.device("cpu")}
    ns.update(injected)
    exec(compile(ast.Module(body=body, type_ignores=[]), rel_path, "exec"), ns)
    return ns


def reference_optimizer(model, fourier: bool):
    """Param groups exactly as CYL:592-612 / FCYL:607-632 (top-level code, restated)."""
    import torch

    groups = [
        {"params": [model.c_coef] + list(model.sin_coef_list) + list(model.cos_coef_list) + list(model.tanh_coef_list),
         "lr": 1e-2},
        {"params": list(model.phi_list), "lr": 1e-3},
        {"params": list(model.z_list) + list(model.zsin_list) + list(model.zcos_list) + list(model.ztanh_list), "lr": 1e-2},
        {"params": list(model.omega_list), "lr": 1e3},
    ]
    if fourier:
        groups.append({"params": list(model.period_list) + list(model.trig_period_list), "lr": 1e-2})
    return torch.optim.Adamax(groups, weight_decay=0.0)


def reference_losses(model, snapshot, beta: float, l1_lambda: float):
    """Loss assembly of CYL:711-733 (top-level code, restated op for op)."""
    import torch

    recon, latent_spatial, _ = model(snapshot)
    ortho = 0
    r = latent_spatial.size(1)
    for i in range(r):
        for j in range(i + 1, r):
            ortho = ortho + torch.norm(latent_spatial[:, i] @ latent_spatial[:, j].T, p="fro")
    mse = torch.nn.MSELoss()(recon, snapshot)
    l1 = torch.norm(model.c_coef, p=1)
    for lst in (model.sin_coef_list, model.cos_coef_list, model.tanh_coef_list):
        for c in lst:
            l1 = l1 + torch.norm(c, p=1)
    return mse, ortho, l1, mse + beta * ortho + l1_lambda * l1
