"""torch custom ops over the C ABI (one op per hot-path entry point of include/desmo_b200.h).

`torch.ops.desmo_b200.build_w / fused_residual_grad / adamax_update` take the packed device tensors of a DesmoEngine and launch the
sm_100a kernels on the current stream.  CUDA only: there is no CPU implementation registered, so calling them with CPU tensors
fails loudly in the dispatcher."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check, make_shape


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.DesmoError("desmo_b200 ops are CUDA-only (sm_100a); no CPU fallback")


@torch.library.custom_op("desmo_b200::build_w", mutates_args=("rows", "W", "step", "workspace"))
def build_w(gates: torch.Tensor, rows: torch.Tensor, coefs: Optional[torch.Tensor], periods: Optional[torch.Tensor], W: torch.Tensor,
            step: torch.Tensor, workspace: torch.Tensor, n: int, n_global: int, m: int, r: int, polyorder: int, nF: int, path: int) -> None:
    _require_cuda(gates, rows, W)
    s = make_shape(n, m, r, polyorder, nF, n_global, path)
    with torch.cuda.device(gates.device):
        check(_lib.load().desmo_build_w(C.byref(s), _p(gates), _p(rows), _p(coefs), _p(periods), _p(W), _p(step), _p(workspace),
                                        _stream(gates)), "desmo_build_w")


@torch.library.custom_op("desmo_b200::fused_residual_grad", mutates_args=("dphi", "red", "workspace"))
def fused_residual_grad(U: torch.Tensor, P: torch.Tensor, phi: torch.Tensor, omega: torch.Tensor, W: torch.Tensor, dphi: torch.Tensor,
                        red: torch.Tensor, workspace: torch.Tensor, n: int, n_global: int, m: int, r: int, polyorder: int, nF: int,
                        path: int) -> None:
    _require_cuda(U, P, phi, omega, W, dphi, red)
    s = make_shape(n, m, r, polyorder, nF, n_global, path)
    with torch.cuda.device(U.device):
        check(_lib.load().desmo_fused_residual_grad(C.byref(s), _p(U), _p(P), _p(phi), _p(omega), _p(W), _p(dphi), _p(red),
                                                    _p(workspace), _stream(U)), "desmo_fused_residual_grad")


def engine_step_via_ops(e) -> None:
    """One train step of a DesmoEngine routed through the registered ops (build_w + fused pass), then the update."""
    torch.ops.desmo_b200.build_w(e.gates, e.rows, e.coefs, e.periods, e.W, e.step_dev, e.workspace, e.n, e.n_global, e.m, e.r,
                                 e.polyorder, e.nF, e.shape.path)
    torch.ops.desmo_b200.fused_residual_grad(e.U, e.P, e.phi, e.omega, e.W, e.dphi, e.red, e.workspace, e.n, e.n_global, e.m, e.r,
                                             e.polyorder, e.nF, e.shape.path)
    e.all_reduce()
    e.adamax_update()
