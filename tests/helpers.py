"""Shared helpers for the parity tests: oracle <-> engine parameter transfer and seeded cases."""
import json
import os

import numpy as np

from oracle import desmo_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make_case(kind, n, m, r, p, nF=None, period_init=60.0, omega_init=10000.0, data_seed=0, perturb_seed=42, perturb_rel=0.1):
    X = orc.synthetic_snapshots(kind, n, m, data_seed)
    modes, _, _, _ = orc.pod_analysis(X, r)
    snap = np.ascontiguousarray(X.T.astype(np.float32))
    base = orc.init_params(n, m, p, r, omega_init=omega_init, nF=nF, period_init=period_init)
    prm = orc.perturb(base, seed=perturb_seed, rel=perturb_rel) if perturb_rel else base
    return X, modes, snap, prm


def golden_case(name):
    fx = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(fx["meta"]))
    X, modes, snap, prm = make_case(meta["kind"], meta["n"], meta["m"], meta["r"], meta["p"], meta["nF"],
                                    meta["period_init"] or 60.0, meta.get("omega_init", 10000.0), meta["data_seed"],
                                    meta["perturb_seed"], meta["perturb_rel"])
    return fx, meta, modes, snap, prm


def load_engine(engine, prm, modes, snap=None):
    """Copies oracle-packed parameters into a DesmoEngine."""
    import torch

    dev = engine.device
    engine.set_pod_modes(modes)
    engine.phi.zero_()
    engine.phi[:, :engine.n] = torch.from_numpy(prm.phi).to(dev)
    engine.gates.copy_(torch.from_numpy(prm.gates).to(dev))
    engine.omega.copy_(torch.from_numpy(prm.omega).to(dev))
    if prm.fourier:
        engine.coefs.copy_(torch.from_numpy(prm.coefs).to(dev))
        engine.periods.copy_(torch.from_numpy(prm.periods).to(dev))
    else:
        engine.rows.zero_()
        engine.rows[:, :engine.m] = torch.from_numpy(prm.zall).to(dev)
    engine.reset_optimizer()
    if snap is not None:
        engine.set_snapshot(torch.from_numpy(snap))


def engine_params(engine):
    """Engine state -> dict of numpy arrays in the oracle's packed layout."""
    out = {"phi": engine.phi[:, :engine.n].cpu().numpy(), "gates": engine.gates.cpu().numpy(), "omega": engine.omega.cpu().numpy()}
    if engine.nF:
        out["coefs"], out["periods"] = engine.coefs.cpu().numpy(), engine.periods.cpu().numpy()
    else:
        out["zall"] = engine.rows[:, :engine.m].cpu().numpy()
    return out
