"""Golden vectors for the DESMO_AE variant, produced by the REFERENCE's own classes.

TEST INFRASTRUCTURE; runs only where /root/reference exists.  ``POOL_DATA``, ``Autoencoder_Linear_Temporal`` and ``SINDyAutoencoder`` are
AST-extracted from ``DESMO_AE/DESMO_Cylinder_AE-Final.py`` (AE:376-455,629-768) and executed on CPU; the optimizer groups (AE:783-809) and
the loss assembly of the loop (AE:849-862) are top-level code and restated here line for line.  Writes ``tests/golden/ae_r2p2.npz``:
inputs, the initial ``state_dict``, forward outputs, the five losses, the gradients of every library parameter and of the first / last
MLP layers, and the state after 20 Adamax steps.

The shipped initial frequencies (omega_phi = 1e4 / 1e3) make sin(omega * code) chaotic in fp32 (a 1e-7 difference in the code flips the
phase), so -- as for the DESMO goldens -- the fixture uses tame frequencies and a tame frequency learning rate through the same code path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as ref  # noqa: E402

AE = "DESMO_AE/DESMO_Cylinder_AE-Final.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
OMEGAS = (3.0, 2.0, 2.5, 1.5, 1.0, 1.0)  # omega_phi1..6
STEPS = 20
OMEGA_LR = 1e-2
KEEP_GRADS = ("temporal_ae.encoder.0.weight", "temporal_ae.encoder.0.bias", "temporal_ae.encoder.12.weight", "temporal_ae.decoder.0.weight",
              "temporal_ae.decoder.12.weight", "temporal_ae.decoder.12.bias")


def snapshot(n: int, m: int, seed: int = 5) -> np.ndarray:
    """(m, n) fp32: a travelling periodic field + noise, temporal mean removed (CYL:136-149 semantics)."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, n)[None, :]
    t = np.arange(m)[:, None]
    u = sum(np.sin(2 * np.pi * (k + 1) * x + 0.3 * k) * np.cos(2 * np.pi * (k + 1) * t / 20.0 + 0.1 * k) / (k + 1) for k in range(4))
    u = u + 0.01 * rng.standard_normal((m, n))
    u = u - u.mean(axis=0, keepdims=True)
    return np.ascontiguousarray(u.astype(np.float32))


def reference_optimizer(model):
    """AE:783-809, restated."""
    f_g = ['c_coef', 'sin_coef_1', 'tanh_coef_1', 'tanh_coef_2', 'ztanh_coef_1', 'ztanh_coef_2', 'cos_coef_1', 'sin_coef_2', 'cos_coef_2',
           'zsin_coef_1', 'zcos_coef_1', 'zsin_coef_2', 'zcos_coef_2']
    om = ['omega_phi1', 'omega_phi2', 'omega_phi3', 'omega_phi4', 'omega_phi5', 'omega_phi6']
    return torch.optim.Adamax([
        {'params': [p for n_, p in model.named_parameters() if n_ in om], 'lr': 1e2},
        {'params': [p for n_, p in model.named_parameters() if n_ in f_g and n_ not in om], 'lr': 1e-2},
        {'params': [p for n_, p in model.named_parameters() if n_ not in f_g and n_ not in om]},
    ], lr=1e-2, weight_decay=0.0)


def reference_losses(model, snap, beta=1e-3, l1_lambda=1e-6, ae_beta=1e-3):
    """AE:849-862, restated op for op (the ortho term is the reference's MSE of the n x n outer product against zeros(1))."""
    criterion = torch.nn.MSELoss()
    recon, latent_spatial, _, ae_rec = model(snap)
    ortho = criterion(latent_spatial[:, 0] @ latent_spatial[:, 1].T, torch.zeros(1, device=snap.device))
    loss = criterion(recon, snap)
    ae_loss = criterion(ae_rec, snap)
    l1 = (torch.norm(model.c_coef, p=1) + torch.norm(model.cos_coef_1, p=1) + torch.norm(model.cos_coef_2, p=1)
          + torch.norm(model.sin_coef_1, p=1) + torch.norm(model.sin_coef_2, p=1))
    total = loss + beta * ortho + l1_lambda * l1 + ae_beta * ae_loss
    return loss, ortho, l1, ae_loss, total, (recon, latent_spatial, ae_rec)


def perturb_(model, seed=44, rel=0.1):
    """Moves the all-ones library parameters off their symmetric start (same recipe as the DESMO goldens)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.startswith("temporal_ae") or name.startswith("omega"):
                continue
            p.mul_(1.0 + rel * torch.randn(p.shape, generator=g))
        for i, w in enumerate(OMEGAS):
            getattr(model, f"omega_phi{i + 1}").fill_(w)


def main():
    import warnings

    warnings.filterwarnings("ignore")
    n, m, p, r = 257, 40, 2, 2
    torch.set_num_threads(4)
    ns = ref.load_definitions(AE, ("POOL_DATA", "binomial_coefficient", "calculate_number_of_terms", "Autoencoder_Linear_Temporal",
                                   "SINDyAutoencoder"), {"r": r, "polyorder": p})
    torch.manual_seed(43)
    model = ns["SINDyAutoencoder"](n, m, p, r)
    perturb_(model)
    snap = torch.from_numpy(snapshot(n, m))
    fx = {"meta": json.dumps(dict(n=n, m=m, polyorder=p, r=r, omegas=list(OMEGAS), steps=STEPS, omega_lr=OMEGA_LR, beta=1e-3, l1_lambda=1e-6, ae_beta=1e-3)),
          "snapshot": snap.numpy(), "keys": np.array(list(model.state_dict().keys()))}
    for k, v in model.state_dict().items():
        fx["init/" + k] = v.detach().numpy().copy()
    loss, ortho, l1, ae_loss, total, (recon, lat, ae_rec) = reference_losses(model, snap)
    model.zero_grad()
    total.backward()
    fx.update(loss=np.float64(loss.item()), ortho=np.float64(ortho.item()), l1=np.float64(l1.item()), ae_loss=np.float64(ae_loss.item()),
              total=np.float64(total.item()), recon=recon.detach().numpy(), latent=lat.detach().numpy(), ae_rec=ae_rec.detach().numpy())
    for name, q in model.named_parameters():
        if name.startswith("temporal_ae") and name not in KEEP_GRADS:
            continue
        fx["grad/" + name] = (q.grad if q.grad is not None else torch.zeros_like(q)).numpy().copy()
        fx["hasgrad/" + name] = np.bool_(q.grad is not None)
    opt = reference_optimizer(model)
    opt.param_groups[0]["lr"] = OMEGA_LR  # the shipped 1e2 moves the frequencies by ~100 per step: chaotic again after one step
    hist = []
    for _ in range(STEPS):
        loss, ortho, l1, ae_loss, total, _o = reference_losses(model, snap)
        opt.zero_grad()
        total.backward()
        opt.step()
        hist.append((loss.item(), ortho.item(), l1.item(), ae_loss.item(), total.item()))
    fx["hist"] = np.array(hist)
    for k, v in model.state_dict().items():
        if k.startswith("temporal_ae") and k not in KEEP_GRADS:
            continue
        fx["final/" + k] = v.detach().numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ae_r2p2.npz"), **fx)
    print("ae_r2p2: loss", hist[0], "->", hist[-1])


if __name__ == "__main__":
    main()
