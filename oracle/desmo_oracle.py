"""CPU oracle for the DESMO training hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``desmo_b200``) never imports it and has no CPU fallback.

This is a numpy *restatement* (closed-form forward + hand-derived backward) of what
the reference computes with torch autograd.  Every function cites the reference
lines it follows (paths relative to ``/root/reference``; ``CYL`` =
``DESMO/cylinder_flow/DESMO-Cylinder.py``, ``FCYL`` =
``DESMO_Fourier/cylinder_flow/DESMO-Cylinder.py``).

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/make_golden.py`` executes the reference's own ``POOL_DATA`` /
``DESMO`` / ``DESMOFourier`` / ``fourier_series`` / ``poly_norm`` /
``nonlinear_norm`` definitions (AST-extracted from the read-only tree, run with
torch autograd + ``torch.optim.Adamax`` on CPU) on seeded inputs and commits the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against them, plus the known-answer facts in the shipped logs/checkpoints
(library sizes, parameter totals, ``l1(init) == K``, state-dict key order).

Packed parameter layout used here and by the CUDA path (K = T + 3r):

    phi    (r, n)      phi_list[i]                                    CYL:506
    gates  (K,)        [c_coef (T) | sin_coef (r) | cos_coef (r) | tanh_coef (r)]   CYL:513,524-526
    zall   (K, m)      [z_list (T) | zsin (r) | zcos (r) | ztanh (r)]  CYL:516-521   (DESMO)
    coefs  (K, 2nF+1)  same row order, Fourier coefficients            FCYL:527,532-534 (DESMOFourier)
    periods(K,)        [period_list (T) | trig periods of sin (r) | cos (r) | tanh (r)]  FCYL:528-529
    omega  (3r,)       REFERENCE order: omega[3i], omega[3i+1], omega[3i+2] = sin, cos, tanh of mode i   CYL:530,561-563

The reference's ``trig_period_list[3i+j]`` maps to ``periods[T + j*r + i]``.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# ----------------------------------------------------------------------------------------------
# library bookkeeping
# ----------------------------------------------------------------------------------------------


def binomial_coefficient(n: int, k: int) -> int:
    """n choose k, 0 when k > n (CYL:440-446)."""
    if k > n:
        return 0
    return math.factorial(n) // (math.factorial(k) * math.factorial(n - k))


def number_of_terms(n_vars: int, polyorder: int) -> int:
    """T = sum_{k<=p} C(r+k-1, k)  (CYL:448-455)."""
    return sum(binomial_coefficient(n_vars + k - 1, k) for k in range(polyorder + 1))


def monomial_table(n_vars: int, polyorder: int) -> List[Tuple[int, ...]]:
    """Index tuples of every library column in the reference's column order.

    POOL_DATA (CYL:376-434) appends the constant column, then for degree d = 1..p the
    nested ``for i; for j in range(i, nVars); ...`` loops, i.e. exactly
    ``itertools.combinations_with_replacement(range(nVars), d)``.
    """
    if not 0 <= polyorder <= 7:
        raise ValueError("POOL_DATA supports polyorder 0..7 (CYL:376-434)")
    table: List[Tuple[int, ...]] = [()]
    for d in range(1, polyorder + 1):
        table.extend(itertools.combinations_with_replacement(range(n_vars), d))
    return table


def pool_data(latent: np.ndarray, polyorder: int) -> np.ndarray:
    """(n, r) -> (n, T) monomial library; products taken left to right as CYL:390-431."""
    n, r = latent.shape
    cols = []
    for idx in monomial_table(r, polyorder):
        col = np.ones(n, dtype=latent.dtype)
        for pos, v in enumerate(idx):
            col = latent[:, v].copy() if pos == 0 else col * latent[:, v]
        cols.append(col)
    return np.stack(cols, axis=1)


def pool_data_derivative(latent: np.ndarray, polyorder: int, d_lib: np.ndarray) -> np.ndarray:
    """sum_j d_lib[:, j] * dL_j/dPhi_i  -> (n, r).  Multiplicity-aware."""
    n, r = latent.shape
    out = np.zeros((n, r), dtype=latent.dtype)
    for j, idx in enumerate(monomial_table(r, polyorder)):
        for pos in range(len(idx)):
            rest = np.ones(n, dtype=latent.dtype)
            for q, v in enumerate(idx):
                if q != pos:
                    rest = rest * latent[:, v]
            out[:, idx[pos]] += d_lib[:, j] * rest
    return out


def term_names(n_vars: int, polyorder: int) -> List[str]:
    """Human-readable names in K order (cf. generate_c_coef_terms, TURB:1252-1259)."""
    names = []
    for idx in monomial_table(n_vars, polyorder):
        names.append("1" if not idx else "*".join(f"phi{v + 1}" for v in idx))
    for fn in ("sin", "cos", "tanh"):
        names.extend(f"{fn}(w*phi{i + 1})" for i in range(n_vars))
    return names


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------


@dataclass
class DesmoParams:
    """Packed trainable state (see module docstring).  ``fourier`` selects DESMOFourier."""

    r: int
    polyorder: int
    n: int
    m: int
    phi: np.ndarray
    gates: np.ndarray
    omega: np.ndarray
    zall: Optional[np.ndarray] = None  # DESMO
    coefs: Optional[np.ndarray] = None  # DESMOFourier
    periods: Optional[np.ndarray] = None  # DESMOFourier
    nF: int = 0

    @property
    def fourier(self) -> bool:
        return self.coefs is not None

    @property
    def T(self) -> int:
        return number_of_terms(self.r, self.polyorder)

    @property
    def K(self) -> int:
        return self.T + 3 * self.r

    def group_arrays(self) -> List[List[str]]:
        """Optimizer param groups in the reference's order (CYL:592-612, FCYL:607-632)."""
        groups = [["gates"], ["phi"], ["coefs" if self.fourier else "zall"], ["omega"]]
        if self.fourier:
            groups.append(["periods"])
        return groups

    def copy(self) -> "DesmoParams":
        kw = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()}
        return DesmoParams(**kw)

    def num_parameters(self) -> int:
        return sum(getattr(self, k).size for g in self.group_arrays() for k in g)


def init_params(n: int, m: int, polyorder: int, r: int, omega_init: float = 10000.0,
                nF: Optional[int] = None, period_init: float = 60.0,
                dtype=np.float32) -> DesmoParams:
    """Constant initialisation of CYL:501-530 / FCYL:513-544 (everything 1, omega = omega_init)."""
    T = number_of_terms(r, polyorder)
    K = T + 3 * r
    p = DesmoParams(r=r, polyorder=polyorder, n=n, m=m,
                    phi=np.ones((r, n), dtype), gates=np.ones(K, dtype),
                    omega=np.full(3 * r, omega_init, dtype))
    if nF is None:
        p.zall = np.ones((K, m), dtype)
    else:
        p.nF = nF
        p.coefs = np.ones((K, 2 * nF + 1), dtype)
        p.periods = np.full(K, period_init, dtype)
    return p


def state_dict_keys(r: int, polyorder: int, fourier: bool) -> List[str]:
    """Key order of ``model.state_dict()`` = nn.Module registration order (CYL:506-530, FCYL:518-544)."""
    T = number_of_terms(r, polyorder)
    keys = ["c_coef"] + [f"phi_list.{i}" for i in range(r)]
    keys += [f"z_list.{j}" for j in range(T)]
    if fourier:
        keys += [f"period_list.{j}" for j in range(T)]
        keys += [f"trig_period_list.{j}" for j in range(3 * r)]
    for name in ("zsin_list", "zcos_list", "ztanh_list", "sin_coef_list", "cos_coef_list", "tanh_coef_list"):
        keys += [f"{name}.{i}" for i in range(r)]
    keys += [f"omega_list.{i}" for i in range(3 * r)]
    return keys


def to_state_dict(p: DesmoParams) -> Dict[str, np.ndarray]:
    """Packed -> reference key layout (shapes as in the shipped .pt files)."""
    T, r = p.T, p.r
    rows = p.coefs if p.fourier else p.zall
    out: Dict[str, np.ndarray] = {"c_coef": p.gates[:T].copy()}
    for i in range(r):
        out[f"phi_list.{i}"] = p.phi[i].copy()
    for j in range(T):
        out[f"z_list.{j}"] = rows[j].copy()
    if p.fourier:
        for j in range(T):
            out[f"period_list.{j}"] = p.periods[j:j + 1].copy()
        for i in range(r):
            for b in range(3):
                out[f"trig_period_list.{3 * i + b}"] = p.periods[T + b * r + i:T + b * r + i + 1].copy()
    for b, name in enumerate(("zsin_list", "zcos_list", "ztanh_list")):
        for i in range(r):
            out[f"{name}.{i}"] = rows[T + b * r + i].copy()
    for b, name in enumerate(("sin_coef_list", "cos_coef_list", "tanh_coef_list")):
        for i in range(r):
            out[f"{name}.{i}"] = np.asarray(p.gates[T + b * r + i]).copy()
    for i in range(3 * r):
        out[f"omega_list.{i}"] = np.asarray(p.omega[i]).copy()
    return {k: out[k] for k in state_dict_keys(r, p.polyorder, p.fourier)}


def from_state_dict(sd: Dict[str, np.ndarray], r: int, polyorder: int, dtype=np.float32) -> DesmoParams:
    """Reference key layout -> packed."""
    fourier = "period_list.0" in sd
    T = number_of_terms(r, polyorder)
    K = T + 3 * r
    g = lambda k: np.asarray(sd[k], dtype=dtype)  # noqa: E731
    n = g("phi_list.0").shape[0]
    width = g("z_list.0").shape[0]
    rows = np.zeros((K, width), dtype)
    gates = np.zeros(K, dtype)
    gates[:T] = g("c_coef")
    for j in range(T):
        rows[j] = g(f"z_list.{j}")
    for b, (zn, cn) in enumerate((("zsin_list", "sin_coef_list"), ("zcos_list", "cos_coef_list"),
                                  ("ztanh_list", "tanh_coef_list"))):
        for i in range(r):
            rows[T + b * r + i] = g(f"{zn}.{i}")
            gates[T + b * r + i] = g(f"{cn}.{i}").reshape(())
    p = DesmoParams(r=r, polyorder=polyorder, n=n, m=width, gates=gates,
                    phi=np.stack([g(f"phi_list.{i}") for i in range(r)]),
                    omega=np.array([g(f"omega_list.{i}").reshape(()) for i in range(3 * r)], dtype))
    if fourier:
        periods = np.zeros(K, dtype)
        for j in range(T):
            periods[j] = g(f"period_list.{j}").reshape(())
        for i in range(r):
            for b in range(3):
                periods[T + b * r + i] = g(f"trig_period_list.{3 * i + b}").reshape(())
        p.coefs, p.periods, p.nF = rows, periods, (width - 1) // 2
        p.m = -1  # not recoverable from a Fourier checkpoint; caller sets it
    else:
        p.zall = rows
    return p


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------


def t_points(m: int, dtype=np.float32) -> np.ndarray:
    """``torch.linspace(0, m, m)`` (FCYL:485): spacing m/(m-1), evaluated like ATen's CPU linspace
    (symmetric fill: start + i*step for the first half, end - (m-1-i)*step for the second)."""
    step = (dtype(m) - dtype(0)) / dtype(m - 1)
    i = np.arange(m)
    lo = (dtype(0) + step * i.astype(dtype)).astype(dtype)
    hi = (dtype(m) - step * (m - 1 - i).astype(dtype)).astype(dtype)
    return np.where(i < m // 2, lo, hi).astype(dtype)


def fourier_angles(m: int, nF: int, periods: np.ndarray, dtype=np.float32) -> np.ndarray:
    """theta[k, h-1, t] = ((2*pi*h) * t) / period_k in the reference's operation order (FCYL:504):
    python-float ``2*torch.pi*n`` rounded to fp32 when it meets the fp32 tensor, fp32 multiply by t,
    fp32 divide by the (1,)-shaped period."""
    t = t_points(m, dtype)
    h = np.arange(1, nF + 1)
    two_pi_h = (2.0 * math.pi * h).astype(dtype)  # python double product, then cast
    num = (two_pi_h[:, None] * t[None, :]).astype(dtype)  # (nF, m)
    return (num[None, :, :] / periods.astype(dtype)[:, None, None]).astype(dtype)


def fourier_rows(coefs: np.ndarray, periods: np.ndarray, m: int) -> np.ndarray:
    """z_k(t) = a0 + sum_h a_h cos(theta) + b_h sin(theta)  (FCYL:487-506); coeff order [a0, a1, b1, a2, b2, ...]."""
    dtype = coefs.dtype.type
    K, width = coefs.shape
    nF = (width - 1) // 2
    th = fourier_angles(m, nF, periods, dtype)
    z = np.repeat(coefs[:, 0:1], m, axis=1).astype(dtype)
    for h in range(1, nF + 1):
        z = z + (coefs[:, 2 * h - 1:2 * h] * np.cos(th[:, h - 1, :]) + coefs[:, 2 * h:2 * h + 1] * np.sin(th[:, h - 1, :]))
    return z.astype(dtype)


def temporal_rows(p: DesmoParams) -> np.ndarray:
    """(K, m) temporal series of every term: free vectors (CYL:550,558-560) or Fourier series (FCYL:563,570-572)."""
    return fourier_rows(p.coefs, p.periods, p.m) if p.fourier else p.zall


def spatial_library(p: DesmoParams, pod_modes: np.ndarray):
    """G = [L(Phi) | sin | cos | tanh] (n, K) and Phi = phi * POD (n, r)   (CYL:538-548,565-567)."""
    dtype = p.phi.dtype
    lat = (p.phi.T * pod_modes[:, :p.r].astype(dtype)).astype(dtype)  # CYL:538-545
    r = p.r
    lib = pool_data(lat, p.polyorder)  # CYL:548
    om = p.omega.astype(dtype)
    s = np.sin(om[0::3][None, :] * lat)  # CYL:565
    c = np.cos(om[1::3][None, :] * lat)  # CYL:566
    h = np.tanh(om[2::3][None, :] * lat)  # CYL:567
    return np.concatenate([lib, s, c, h], axis=1).astype(dtype), lat


def forward(p: DesmoParams, pod_modes: np.ndarray):
    """Returns (recon (m, n), latent_spatial (n, r), z_values (T, m)) like CYL:576."""
    G, lat = spatial_library(p, pod_modes)
    zrows = temporal_rows(p)
    W = p.gates[:, None] * zrows
    recon = G @ W  # CYL:572  (n, m)
    return recon.T, lat, zrows[:p.T]


# ----------------------------------------------------------------------------------------------
# loss + closed-form gradients
# ----------------------------------------------------------------------------------------------


@dataclass
class StepOutput:
    mse: float
    ortho: float
    l1: float
    total: float
    grads: Dict[str, np.ndarray] = field(default_factory=dict)
    E: Optional[np.ndarray] = None  # G^T R  (K, m), unscaled
    gram: Optional[np.ndarray] = None  # Phi^T Phi (r, r)


def loss_and_grads(p: DesmoParams, pod_modes: np.ndarray, snapshot: np.ndarray, beta: float,
                   l1_lambda: float, want_grads: bool = True) -> StepOutput:
    """Loss of CYL:714-733 and its gradient w.r.t. every packed parameter (what ``total_loss.backward()``
    CYL:766 produces).  ``snapshot`` is the reference's (m, n) batch (CYL:708)."""
    dtype = p.phi.dtype
    n, m, r, T = p.n, p.m, p.r, p.T
    assert snapshot.shape == (m, n)
    G, lat = spatial_library(p, pod_modes)
    zrows = temporal_rows(p)
    W = (p.gates[:, None] * zrows).astype(dtype)
    R = (G @ W - snapshot.T.astype(dtype)).astype(dtype)  # (n, m)
    mse = float(np.mean(R.astype(np.float64) ** 2)) if dtype == np.float64 else float(np.mean(R * R, dtype=np.float32))
    gram = (lat.T @ lat).astype(dtype)
    ortho = float(sum(abs(gram[i, j]) for i in range(r) for j in range(i + 1, r)))  # CYL:714-720
    l1 = float(np.sum(np.abs(p.gates)))  # CYL:725-731
    out = StepOutput(mse=mse, ortho=ortho, l1=l1, total=mse + beta * ortho + l1_lambda * l1, gram=gram)
    if not want_grads:
        return out

    scale = dtype.type(2.0 / (n * m))
    Eraw = (G.T @ R).astype(dtype)  # (K, m)
    E = scale * Eraw
    D = scale * (R @ W.T)  # (n, K)
    out.E = Eraw
    grads: Dict[str, np.ndarray] = {}
    dz = p.gates[:, None] * E  # d total / d temporal rows
    grads["gates"] = (np.sum(zrows * E, axis=1) + l1_lambda * np.sign(p.gates)).astype(dtype)
    om = p.omega.astype(dtype)
    a_s, a_c, a_h = om[0::3][None, :] * lat, om[1::3][None, :] * lat, om[2::3][None, :] * lat
    cs, sn, th = np.cos(a_s), np.sin(a_c), np.tanh(a_h)
    Ds, Dc, Dh = D[:, T:T + r], D[:, T + r:T + 2 * r], D[:, T + 2 * r:T + 3 * r]
    sech2 = 1.0 - th * th
    dlat = pool_data_derivative(lat, p.polyorder, D[:, :T])
    dlat = dlat + Ds * om[0::3][None, :] * cs - Dc * om[1::3][None, :] * sn + Dh * om[2::3][None, :] * sech2
    sgn = np.sign(gram)
    np.fill_diagonal(sgn, 0.0)
    dlat = dlat + dtype.type(beta) * (lat @ sgn.T)  # d/dPhi_i sum_{i<j}|Phi_i.Phi_j| = sum_{j!=i} sign(dot_ij) Phi_j
    grads["phi"] = (dlat * pod_modes[:, :r].astype(dtype)).T.astype(dtype)
    domega = np.zeros(3 * r, dtype)
    domega[0::3] = np.sum(Ds * lat * cs, axis=0)
    domega[1::3] = -np.sum(Dc * lat * sn, axis=0)
    domega[2::3] = np.sum(Dh * lat * sech2, axis=0)
    grads["omega"] = domega
    if p.fourier:
        nF = p.nF
        thg = fourier_angles(m, nF, p.periods, dtype.type)  # (K, nF, m)
        dco = np.zeros_like(p.coefs)
        dco[:, 0] = np.sum(dz, axis=1)
        dper = np.zeros_like(p.periods)
        for h in range(1, nF + 1):
            c_, s_ = np.cos(thg[:, h - 1, :]), np.sin(thg[:, h - 1, :])
            dco[:, 2 * h - 1] = np.sum(dz * c_, axis=1)
            dco[:, 2 * h] = np.sum(dz * s_, axis=1)
            # d theta / d period = -theta / period
            dper += np.sum(dz * (thg[:, h - 1, :] / p.periods[:, None]) *
                           (p.coefs[:, 2 * h - 1:2 * h] * s_ - p.coefs[:, 2 * h:2 * h + 1] * c_), axis=1)
        grads["coefs"], grads["periods"] = dco.astype(dtype), dper.astype(dtype)
    else:
        grads["zall"] = dz.astype(dtype)
    out.grads = grads
    return out


# ----------------------------------------------------------------------------------------------
# optimizer / scheduler (torch semantics restated)
# ----------------------------------------------------------------------------------------------

REFERENCE_LRS = (1e-2, 1e-3, 1e-2, 1e3, 1e-2)  # gates, phi, z, omega, (periods)  CYL:592-612, FCYL:607-632


class Adamax:
    """torch.optim.Adamax (torch 2.11 ``_single_tensor_adamax``): exp_avg.lerp_(g, 1-b1);
    exp_inf = max(b2*exp_inf, |g|+eps); p -= (lr / (1-b1^t)) * exp_avg / exp_inf.  weight_decay = 0 (CYL:612)."""

    def __init__(self, params: DesmoParams, lrs=REFERENCE_LRS, betas=(0.9, 0.999), eps=1e-8):
        self.groups = params.group_arrays()
        self.lrs = [float(lrs[i]) for i in range(len(self.groups))]
        self.b1, self.b2, self.eps = betas[0], betas[1], eps
        self.t = 0
        self.exp_avg = {k: np.zeros_like(getattr(params, k)) for g in self.groups for k in g}
        self.exp_inf = {k: np.zeros_like(getattr(params, k)) for g in self.groups for k in g}

    def step(self, params: DesmoParams, grads: Dict[str, np.ndarray]) -> None:
        self.t += 1
        bc = 1.0 - self.b1 ** self.t
        for lr, names in zip(self.lrs, self.groups):
            for k in names:
                prm = getattr(params, k)
                dt = prm.dtype.type
                g = grads[k].astype(prm.dtype)
                m_, u_ = self.exp_avg[k], self.exp_inf[k]
                m_ += dt(1.0 - self.b1) * (g - m_)
                np.maximum(u_ * dt(self.b2), np.abs(g) + dt(self.eps), out=u_)
                prm += dt(-(lr / bc)) * m_ / u_


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', factor=0.1, threshold=1e-4 rel, cooldown=0,
    min_lr=1e-6, eps=1e-8) as configured at CYL:614."""

    def __init__(self, opt: Adamax, patience: int, factor=0.1, min_lr=1e-6, threshold=1e-4, eps=1e-8):
        self.opt, self.patience, self.factor, self.min_lr = opt, patience, factor, min_lr
        self.threshold, self.eps = threshold, eps
        self.best, self.bad = math.inf, 0

    def step(self, metric: float) -> None:
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            for i, old in enumerate(self.opt.lrs):
                new = max(old * self.factor, self.min_lr)
                if old - new > self.eps:
                    self.opt.lrs[i] = new
            self.bad = 0


def train(p: DesmoParams, pod_modes: np.ndarray, snapshot: np.ndarray, steps: int, beta: float, l1_lambda: float,
          patience: int = 1000, sched_every: int = 10, lrs=REFERENCE_LRS, record_every: int = 1):
    """The hot loop of CYL:706-778 (one full batch per epoch; scheduler every ``sched_every`` epochs on the
    total loss, CYL:776-778; TURB:672 / ANEU:613 use 1).  Mutates ``p``; returns the per-step loss history."""
    opt = Adamax(p, lrs)
    sch = ReduceLROnPlateau(opt, patience)
    hist = []
    for ep in range(steps):
        o = loss_and_grads(p, pod_modes, snapshot, beta, l1_lambda)
        opt.step(p, o.grads)
        if ep % record_every == 0:
            hist.append((o.mse, o.ortho, o.l1, o.total))
        if ep % sched_every == 0:
            sch.step(o.total)
    return np.array(hist), opt, sch


# ----------------------------------------------------------------------------------------------
# POD and post-hoc sparsification
# ----------------------------------------------------------------------------------------------


def pod_analysis(X: np.ndarray, r: int):
    """fp64 thin SVD, modes = U[:, :r], relative error of the rank-r reconstruction (CYL:197-211)."""
    U, S, Vt = np.linalg.svd(X, full_matrices=False)
    modes = U[:, :r]
    approx = modes @ np.diag(S[:r]) @ Vt[:r, :]
    err = np.linalg.norm(X - approx) / np.linalg.norm(X)
    return modes, Vt[:r, :], S, err


def term_norms(p: DesmoParams, pod_modes: Optional[np.ndarray] = None, physical: bool = False) -> np.ndarray:
    """Per-term norms in K order as the reference's post-hoc sweep computes them (poly_norm CYL:624-647, nonlinear_norm
    CYL:653-692; Fourier scripts FCYL:644-720), closed form of ``torch.norm(gate * (lib_col @ z.T))``.

    Reference semantics (default), reproduced on purpose because the active mask is defined by them:
      * the scripts pass the RAW ``model_desmo.phi_list`` (CYL:1192-1194, FCYL:1195-1197), so the library is evaluated on
        phi alone, NOT on phi * POD_modes as in forward();
      * DESMO: ``zs = stack(z_list, dim=1)`` is (m, T) and ``zs[:, i:i+1]`` is term i's series: norm_i = |c_i| ||L_i|| ||z_i||;
      * DESMOFourier: ``zs = stack(..., dim=0)`` is (T, m) but is sliced the same way (FCYL:652,659), so ``zs[:, i:i+1]`` holds
        ALL T polynomial series at time index i: norm_i = |c_i| ||L_i|| sqrt(sum_j z_j(t_i)^2).  The sin/cos/tanh norms use
        the term's own series in both variants (FCYL:693-704).
    ``physical=True`` gives the Frobenius norm of the term as it enters forward() (library of phi * POD, own series)."""
    if physical:
        G, _ = spatial_library(p, pod_modes)
    else:
        ones = np.ones((p.n, p.r), dtype=p.phi.dtype)
        G, _ = spatial_library(p, ones)
    z = temporal_rows(p).astype(np.float64)
    zn = np.linalg.norm(z, axis=1)
    if p.fourier and not physical:
        T = p.T
        if T > p.m:
            raise ValueError("the Fourier scripts' poly_norm indexes time step i for term i: needs T <= m")
        zn[:T] = np.sqrt(np.sum(z[:T, :T] ** 2, axis=0))  # column i of the (T, m) stack
    return np.abs(p.gates.astype(np.float64)) * np.linalg.norm(G.astype(np.float64), axis=0) * zn


def active_mask(norms: np.ndarray, gates: np.ndarray, threshold: float) -> np.ndarray:
    """Gates whose term norm is < threshold are zeroed (CYL:1228-1238); mask = surviving non-zero gates (CYL:1260-1265)."""
    return (norms >= threshold) & (gates != 0)


def relative_error(p: DesmoParams, pod_modes: np.ndarray, snapshot: np.ndarray, mask: Optional[np.ndarray] = None) -> float:
    """||X - recon^T|| / ||X|| with masked gates (CYL:1240-1257)."""
    q = p.copy()
    if mask is not None:
        q.gates = np.where(mask, q.gates, 0).astype(q.gates.dtype)
    recon, _, _ = forward(q, pod_modes)
    return float(np.linalg.norm(snapshot.astype(np.float64) - recon) / np.linalg.norm(snapshot.astype(np.float64)))


def removal_order(norms: np.ndarray, T: int, r: int) -> List[int]:
    """Order in which the greedy sweep removes terms (packed K indices).  The reference lists the polynomial terms first, then
    (sin_i, cos_i, tanh_i) for i = 0..r-1 (TURB:1173-1181) and sorts that list by norm with Python's stable sort (TURB:1183)."""
    listed = list(range(T)) + [T + b * r + i for i in range(r) for b in range(3)]
    return sorted(listed, key=lambda k: float(norms[k]))


def greedy_removal(p: DesmoParams, pod_modes: np.ndarray, snapshot: np.ndarray):
    """Greedy term-removal sweep (TURB:1166-1245): for step = 0..K zero the gates of the ``step`` smallest-norm terms, evaluate
    ||X - recon^T|| / ||X|| (TURB:1227) and count the non-zero gates left (TURB:1229-1234).  Returns [(step, error, nonzero)]."""
    order = removal_order(term_norms(p), p.T, p.r)
    out = []
    for step in range(p.K + 1):
        mask = np.ones(p.K, bool)
        mask[order[:step]] = False
        out.append((step, relative_error(p, pod_modes, snapshot, mask), int(np.count_nonzero(np.where(mask, p.gates, 0)))))
    return out


# ----------------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so both sides see identical data
# ----------------------------------------------------------------------------------------------


def subtract_mean(X: np.ndarray) -> np.ndarray:
    """Remove the temporal mean of every row (CYL:136-149)."""
    return X - X.mean(axis=1, keepdims=True)


def preprocess(X_raw: np.ndarray, d_in: int, d_use: int, magnitude: bool = True, subtract: bool = True, scale_sqrt_m: bool = False,
               t_stride: int = 1):
    """The reference's numpy pre-processing between the reader and POD/training, float64 throughout.

    X_raw is the reader's matrix (n*d_in rows x m_in snapshots, the d_in components of a point on adjacent rows, CYL:52-85).
    convert3Dto2D_data drops every third row (CYL:88-106; d_in=3, d_use=2); convertToMagnitude takes sqrt(sum(square)) over
    the d_use components of each point (CYL:109-133); subtract_mean removes the temporal mean of every row (CYL:136-149) and,
    in the aneurysm scripts, multiplies by 1/sqrt(m) (ANEU:130-144); the channel script then keeps every second snapshot
    (TURB:189, after the mean).  Returns (X, X_mean) with X (n x ceil(m_in/t_stride)) float64 -- the training snapshot is
    X.T cast to float32 (CYL:356,708)."""
    X = np.asarray(X_raw, np.float64)
    if d_use != d_in:
        keep = [i for i in range(X.shape[0]) if i % d_in < d_use]
        X = X[keep]
    if magnitude:
        n = X.shape[0] // d_use
        X = np.sqrt(np.sum(np.square(X.reshape(n, d_use, X.shape[1])), axis=1))
    elif d_use != 1:
        raise ValueError("without magnitude every component is its own row: pass d_in = d_use = 1")
    m_in = X.shape[1]
    X_mean = np.mean(X, 1) if subtract else np.zeros(X.shape[0])
    X = X - X_mean[:, None]
    if scale_sqrt_m:
        X = (1 / np.sqrt(m_in)) * X
    return np.ascontiguousarray(X[:, 0::t_stride]), X_mean


def synthetic_snapshots(kind: str, n: int, m: int, seed: int = 0) -> np.ndarray:
    """fp64 (n, m) data matrix X with the reference's pre-processing applied.

    cylinder : periodic shedding-like field, period ~60 steps (C1/C2)
    channel  : non-periodic random-phase travelling waves, k^-5/3 spectrum (C3)
    aneurysm : pulsatile, 8 smooth spatial fields x 5 harmonics, scaled by 1/sqrt(m) (C4, ANEU:143)
    """
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, n)[:, None]
    t = np.arange(m, dtype=np.float64)[None, :]
    if kind == "cylinder":
        X = sum(np.sin(2 * np.pi * (k + 1) * x + 0.3 * k) * np.cos(2 * np.pi * (k + 1) * t / 60.0 + 0.1 * k) / (k + 1)
                for k in range(6))
        X = X + 0.01 * rng.standard_normal((n, m))
    elif kind == "channel":
        X = np.zeros((n, m))
        for _ in range(40):
            kx = rng.integers(1, 33)
            amp = kx ** (-5.0 / 6.0)
            X += amp * np.sin(2 * np.pi * kx * x + rng.uniform(0, 2 * np.pi) - rng.uniform(0.002, 0.05) * kx * t)
        X += 0.2 * x * (t / m) + 0.05 * rng.standard_normal((n, m))
    elif kind == "aneurysm":
        X = np.zeros((n, m))
        for _ in range(8):
            g = sum(rng.standard_normal() * np.sin(np.pi * (q + 1) * x + rng.uniform(0, np.pi)) / (q + 1) for q in range(4))
            a = rng.standard_normal() + sum(rng.standard_normal() / h * np.cos(2 * np.pi * h * t / m + rng.uniform(0, 2 * np.pi))
                                            for h in range(1, 6))
            X += g * a
        X = X / np.sqrt(m)
    else:
        raise ValueError(kind)
    X = subtract_mean(X)
    return X


def perturb(p: DesmoParams, seed: int = 42, rel: float = 0.1) -> DesmoParams:
    """param *= 1 + rel*N(0,1): moves away from the all-ones init so that gradients w.r.t. every parameter are
    generic and the ortho-term signs are well defined (SURVEY.md section 7 hard parts)."""
    rng = np.random.default_rng(seed)
    q = p.copy()
    for g in q.group_arrays():
        for k in g:
            a = getattr(q, k)
            a *= (1.0 + rel * rng.standard_normal(a.shape)).astype(a.dtype)
    return q
