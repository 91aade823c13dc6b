"""Drop-in ``SINDyAutoencoder`` (the DESMO_AE variant, reference: ``DESMO_AE/DESMO_Cylinder_AE-Final.py`` = ``AE``) on the CUDA engine.

The AE variant replaces ``Phi = phi_list * POD_modes`` by the 2-d code of a temporal auto-encoder: an MLP ``m -> 256 -> ... -> 2`` applied
to every mesh point's time series (``AE:629-683``); everything downstream is DESMO's library model with r = 2 hard-wired
(``AE:688-768``)::

    latent, ae_rec = temporal_ae(X.T)                                   # (n, 2), (n, m)                       AE:745
    recon = (c_coef * POOL_DATA(latent)) @ z_values                     # polynomial library                   AE:752-754,766
          + sin_coef_1 zsin_1 sin(w1 phi1) + cos_coef_1 zcos_1 cos(w2 phi1) + sin_coef_2 zsin_2 sin(w3 phi2) + cos_coef_2 zcos_2 cos(w4 phi2)
    (the tanh terms are built but left out of the sum, AE:761-763: their parameters exist, get no gradient and never move)

Here the MLP stays what it is in the reference -- ``torch.nn.Linear`` layers, i.e. plain library GEMMs -- and the library model runs on
the same fused kernels as ``DESMO``: the engine's ``phi`` buffer takes the encoder's code (POD modes := 1), ``recon`` is produced by
``desmo_reconstruct`` and its backward (``desmo_recon_backward`` + ``desmo_assemble_grads``) returns the gradient of the code to autograd,
which carries it into the encoder.  Parameter names, shapes and registration order are the reference's, so its optimizer-group
construction by name (``AE:783-809``), its loss assembly (``AE:849-862``) and ``state_dict`` work unchanged.  The parameters that enter
the reconstruction are views of the engine's packed buffers (gates / rows / omega in the engine's [sin_i | cos_i | tanh_i] order); the
six tanh-related parameters are ordinary tensors.  No CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .engine import DesmoEngine


class Autoencoder_Linear_Temporal(nn.Module):
    """``AE:629-683``: the same layer stack, initialisation and forward signature (state-dict keys ``encoder.N.weight`` ...)."""

    def __init__(self, m: int):
        super().__init__()
        enc = [m, 256, 128, 64, 16, 8, 4, 2]
        dec = [2, 4, 8, 16, 64, 128, 256, m]

        def stack(w):
            layers = []
            for i in range(len(w) - 1):
                layers.append(nn.Linear(w[i], w[i + 1]))
                if i + 2 < len(w):
                    layers.append(nn.ReLU())
            return nn.Sequential(*layers)

        self.encoder, self.decoder = stack(enc), stack(dec)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(module):
        if isinstance(module, nn.Linear):
            torch.nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                torch.nn.init.zeros_(module.bias)

    def forward(self, x, encoded_input: bool = False):
        if encoded_input:
            return self.decoder(x)
        encoded = self.encoder(x)
        return encoded, self.decoder(encoded)


def reference_state_dict_keys(T: int):
    """Key order of the reference's ``SINDyAutoencoder.state_dict()`` (direct parameters in registration order AE:703-737, then the
    sub-modules ``temporal_ae`` and ``z_list`` in theirs) -- what this module registers; used to pin the layout without a GPU."""
    direct = (["c_coef", "zcos_coef_1", "zcos_coef_2", "zsin_coef_1", "zsin_coef_2", "ztanh_coef_1", "ztanh_coef_2", "cos_coef_1", "cos_coef_2",
               "sin_coef_1", "sin_coef_2", "tanh_coef_1", "tanh_coef_2"] + [f"omega_phi{i}" for i in range(1, 7)])
    mlp = [f"temporal_ae.{part}.{2 * i}.{w}" for part in ("encoder", "decoder") for i in range(7) for w in ("weight", "bias")]
    return direct + mlp + [f"z_list.{j}" for j in range(T)]


class _FusedReconFromCode(torch.autograd.Function):
    """recon (m, n) of the library model for a spatial code ``latent`` (n, 2) coming from upstream autograd (the encoder)."""

    @staticmethod
    def forward(ctx, engine: DesmoEngine, slots, latent, *params):
        with torch.no_grad():
            engine.phi[:, :engine.n].copy_(latent.t())
        ctx.engine, ctx.slots = engine, slots
        return engine.reconstruct()

    @staticmethod
    def backward(ctx, grad_recon):
        eng = ctx.engine
        g = eng.recon_backward(grad_recon.contiguous())
        out = []
        for buf, idx in ctx.slots:
            out.append(g[buf][idx])
        return (None, None, g["phi"].t().contiguous(), *out)  # POD modes are 1: d/d code = d/d phi


class SINDyAutoencoder(nn.Module):
    """``AE:688-768``.  ``forward(X)`` -> ``(recon (m, n), latent_spatial (n, 2, 1), z_values (T, m), ae_rec (m, n))`` with X the (m, n)
    snapshot batch; all four outputs are differentiable as in the reference."""

    def __init__(self, n: int, m: int, polyorder: int, r: int = 2, *, device=None, path: int = _lib.PATH_AUTO):
        super().__init__()
        if r != 2:
            raise ValueError("the reference's SINDyAutoencoder is written for r = 2 (phi1, phi2; AE:746-748)")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.engine = DesmoEngine(n, m, polyorder, r, omega_init=1.0, device=dev, path=path)
        e = self.engine
        T = e.T
        with torch.no_grad():
            e.P.zero_()
            e.P[:, :n] = 1.0                       # Phi = code * 1
            e.gates[T + 2 * r:].zero_()            # tanh terms are not part of the reconstruction (AE:763)
            e.omega.copy_(torch.tensor([10000.0, 1000.0, 1.0, 10000.0, 1000.0, 1.0], device=dev))  # [w1, w2, -, w3, w4, -]  AE:732-737
        P = lambda t: nn.Parameter(t)  # noqa: E731   (aliases the packed storage)
        rows = e.rows[:, :m]
        # registration order == AE:696-737 -> identical state_dict key order
        self.temporal_ae = Autoencoder_Linear_Temporal(m).to(dev)
        print('Number of terms in polynomial library:', T)  # AE:700
        self.c_coef = P(e.gates[:T])
        self.z_list = nn.ParameterList([P(rows[j]) for j in range(T)])
        self.zcos_coef_1, self.zcos_coef_2 = P(rows[T + r + 0]), P(rows[T + r + 1])
        self.zsin_coef_1, self.zsin_coef_2 = P(rows[T + 0]), P(rows[T + 1])
        self.ztanh_coef_1 = nn.Parameter(torch.ones(m, device=dev))
        self.ztanh_coef_2 = nn.Parameter(torch.ones(m, device=dev))
        self.cos_coef_1, self.cos_coef_2 = P(e.gates[T + r + 0]), P(e.gates[T + r + 1])
        self.sin_coef_1, self.sin_coef_2 = P(e.gates[T + 0]), P(e.gates[T + 1])
        self.tanh_coef_1 = nn.Parameter(torch.tensor(1.0, device=dev))
        self.tanh_coef_2 = nn.Parameter(torch.tensor(1.0, device=dev))
        # the reference's naming: omega_phi1 / 2 drive sin / cos of phi1, omega_phi3 / 4 sin / cos of phi2, omega_phi5 / 6 the unused
        # tanh terms (AE:732-737,757-762); the engine's omega is [sin, cos, tanh] per mode
        self.omega_phi1, self.omega_phi2 = P(e.omega[0]), P(e.omega[1])
        self.omega_phi3, self.omega_phi4 = P(e.omega[3]), P(e.omega[4])
        self.omega_phi5 = nn.Parameter(torch.tensor(100.0, device=dev))
        self.omega_phi6 = nn.Parameter(torch.tensor(100.0, device=dev))
        with torch.no_grad():  # AE:732-735 initial values, in the reference's naming
            self.omega_phi1.fill_(10000.0); self.omega_phi2.fill_(1000.0); self.omega_phi3.fill_(10000.0); self.omega_phi4.fill_(1000.0)

    def _views(self):
        e = self.engine
        T, r = e.T, e.r
        rows = e.rows[:, :e.m]
        yield self.c_coef, e.gates[:T], ("gates", slice(0, T))
        for j in range(T):
            yield self.z_list[j], rows[j], ("rows", j)
        for p, k in ((self.zsin_coef_1, T), (self.zsin_coef_2, T + 1), (self.zcos_coef_1, T + r), (self.zcos_coef_2, T + r + 1)):
            yield p, rows[k], ("rows", k)
        for p, k in ((self.sin_coef_1, T), (self.sin_coef_2, T + 1), (self.cos_coef_1, T + r), (self.cos_coef_2, T + r + 1)):
            yield p, e.gates[k], ("gates", k)
        for p, k in ((self.omega_phi1, 0), (self.omega_phi2, 1), (self.omega_phi3, 3), (self.omega_phi4, 4)):
            yield p, e.omega[k], ("omega", k)

    def sync_parameters(self) -> int:
        """Re-aliases parameters user code rebound with ``param.data = ...`` (same contract as ``DESMO.sync_parameters``)."""
        fixed = 0
        for p, v, _ in self._views():
            if p.data_ptr() == v.data_ptr() and p.shape == v.shape:
                continue
            if p.numel() != v.numel() or p.device != v.device or p.dtype != v.dtype:
                raise _lib.DesmoError(f"parameter rebound to an incompatible tensor: {tuple(p.shape)} {p.dtype} {p.device}")
            with torch.no_grad():
                v.copy_(p.data.reshape(v.shape))
            p.data = v
            fixed += 1
        with torch.no_grad():  # the engine's tanh slots stay out of the model whatever a loaded checkpoint says
            self.engine.gates[self.engine.T + 2 * self.engine.r:].zero_()
        return fixed

    def _apply(self, fn, recurse=True):
        before = self.engine.gates.data_ptr()
        out = super()._apply(fn, recurse)
        if self.c_coef.data_ptr() != before:
            raise _lib.DesmoError("desmo_b200 parameters are views of packed CUDA buffers: construct the module with device=... "
                                  "instead of moving / casting it (no CPU fallback)")
        return out

    def forward(self, X):
        e = self.engine
        self.sync_parameters()
        latent, ae_rec = self.temporal_ae(X.T)                    # AE:745   (n, 2), (n, m)
        phi1, phi2 = latent[:, 0].unsqueeze(1), latent[:, 1].unsqueeze(1)
        latent_spatial = torch.stack([phi1, phi2], dim=1)         # AE:748   (n, 2, 1)
        views = list(self._views())
        params = [p for p, _, _ in views]
        slots = [s for _, _, s in views]
        if torch.is_grad_enabled() and (latent.requires_grad or any(p.requires_grad for p in params)):
            recon = _FusedReconFromCode.apply(e, slots, latent, *params)
        else:
            with torch.no_grad():
                e.phi[:, :e.n].copy_(latent.t())
            recon = e.reconstruct()
        z_values = torch.stack(list(self.z_list), dim=0)          # AE:754
        return recon, latent_spatial, z_values, ae_rec.T
