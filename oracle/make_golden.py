"""Generates tests/golden/* by running the REFERENCE'S OWN code (see oracle/reference_loader.py).

Run here (where /root/reference exists):   python oracle/make_golden.py
The outputs are small, committed, and are what pins the oracle (and through it the CUDA path) to the reference.
Inputs are regenerated deterministically inside the tests from the seeds stored in each fixture.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import desmo_oracle as orc  # noqa: E402
from oracle import reference_loader as ref  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, data kind, n, m, r, p, nF, period_init, beta, l1_lambda)
GRAD_CASES = [
    ("cyl_r4p3", "cylinder", 211, 48, 4, 3, None, None, 1e-3, 1e-4),
    ("chan_r4p2", "channel", 160, 40, 4, 2, None, None, 1e-6, 1e-4),
    ("aneu_r3p4", "aneurysm", 130, 36, 3, 4, None, None, 1e-3, 1e-4),
    ("cyl_r2p7", "cylinder", 77, 24, 2, 7, None, None, 1e-3, 1e-4),
    ("fcyl_r2p2", "cylinder", 150, 64, 2, 2, 5, 60.0, 1e-3, 1e-4),
    ("faneu_r2p2", "aneurysm", 120, 50, 2, 2, 7, 1000.0, 1e-3, 1e-4),
]
# The reference's shipped hyper-parameters (omega_init = 1e4, omega lr = 1e3, CYL:584,608) make the training map chaotic:
# a 1e-7 relative difference in d_omega grows ~10x per step (measured: oracle vs reference agree to 2e-7 at step 1,
# 1e-5 at step 4, O(1) by step 8) because omega jumps by ~lr = 1e3 per step and omega*phi wraps many times.  So the
# shipped configuration is pinned over its first 3 steps only, and the 1000-step trajectory gate of north_star is
# pinned on a non-chaotic setting of the SAME code (ctor argument omega_init = 10, omega lr = 1e-2).
# (name, kind, n, m, r, p, nF, period_init, beta, l1_lambda, steps, omega_init, lrs, snapshot steps)
TAME_LRS = (1e-2, 1e-3, 1e-2, 1e-2, 1e-2)
TRAJ_CASES = [
    ("traj_cyl_r4p3", "cylinder", 180, 60, 4, 3, None, None, 1e-3, 1e-4, 1000, 10.0, TAME_LRS, (1, 10, 100, 1000)),
    ("traj_fcyl_r2p2", "cylinder", 180, 60, 2, 2, 6, 60.0, 1e-3, 1e-4, 1000, 10.0, TAME_LRS, (1, 10, 100, 1000)),
    ("traj_default_cyl_r4p3", "cylinder", 180, 60, 4, 3, None, None, 1e-3, 1e-4, 3, 10000.0, orc.REFERENCE_LRS, (1, 2, 3)),
    # the headline library (r = 4, p = 2, K = 27: the tcgen05 path) on channel-like data, two time slabs and three point tiles
    ("traj_chan_r4p2", "channel", 300, 150, 4, 2, None, None, 1e-6, 1e-4, 1000, 10.0, TAME_LRS, (1, 10, 100, 1000)),
]
THRESHOLDS = [float(pow(10, -i)) for i in np.arange(4, -3, -0.5)]  # CYL:1213


def build_inputs(kind, n, m, r, seed=0):
    X = orc.synthetic_snapshots(kind, n, m, seed)
    modes, _, _, _ = orc.pod_analysis(X, r)
    return X, modes


def make_model(r, p, n, m, nF, period_init, pod_modes):
    fourier = nF is not None
    inj = {"POD_modes": pod_modes, "r_DESMO": r, "polyorder": p}
    if fourier:
        inj.update({"t_points": torch.linspace(0, m, m), "period_init": period_init})
        ns = ref.load_definitions(ref.FCYL, ref.FOURIER_DEFS, inj)
        model = ns["DESMOFourier"](n, m, p, r, 10000, nF)
    else:
        ns = ref.load_definitions(ref.CYL, ref.DESMO_DEFS, inj)
        model = ns["DESMO"](n, m, p, r, 10000)
    return model, ns


def load_packed(model, params):
    sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in orc.to_state_dict(params).items()}
    model.load_state_dict(sd, strict=True)


def grads_packed(model, params):
    """Reference autograd grads re-packed into the oracle layout."""
    g = {k: (v.grad.detach().numpy().copy()) for k, v in model.named_parameters()}
    q = orc.from_state_dict(g, params.r, params.polyorder)
    out = {"gates": q.gates, "phi": q.phi, "omega": q.omega}
    if params.fourier:
        out["coefs"], out["periods"] = q.coefs, q.periods
    else:
        out["zall"] = q.zall
    return out


def reference_norms(model, ns, fourier):
    """polynorms / nlnorms exactly as the scripts call them (CYL:1192-1194, FCYL:1195-1197): with the RAW phi_list."""
    with torch.no_grad():
        if fourier:
            pn = ns["poly_norm"](model.c_coef, model.z_list, model.phi_list, model.period_list)
            nl = ns["nonlinear_norm"](model.sin_coef_list, model.cos_coef_list, model.tanh_coef_list, model.zsin_list,
                                      model.zcos_list, model.ztanh_list, model.phi_list, model.omega_list, model.trig_period_list)
        else:
            pn = ns["poly_norm"](model.c_coef, model.z_list, model.phi_list)
            nl = ns["nonlinear_norm"](model.sin_coef_list, model.cos_coef_list, model.tanh_coef_list, model.zsin_list,
                                      model.zcos_list, model.ztanh_list, model.phi_list, model.omega_list)
    return torch.stack(pn), torch.stack(nl)


def reference_threshold_sweep(model, ns, X, snap, fourier):
    """The post-hoc sweep of CYL:1184-1265 / FCYL:1187-1268 run on the reference module (statement for statement; the
    DataLoader round trip is the single full batch ``snap``).  Returns norms and, per threshold, the relative error
    (CYL:1257), the non-zero count (CYL:1260-1265) and the surviving-gate mask in packed K order."""
    model_desmo = model
    original_c_coef = model_desmo.c_coef.clone()
    original_sin_coef_list = [c.clone() for c in model_desmo.sin_coef_list]
    original_cos_coef_list = [c.clone() for c in model_desmo.cos_coef_list]
    original_tanh_coef_list = [c.clone() for c in model_desmo.tanh_coef_list]
    polynorms, nlnorms = reference_norms(model_desmo, ns, fourier)
    errs, counts, masks = [], [], []
    for threshold in THRESHOLDS:
        model_desmo.c_coef.data = original_c_coef.clone()
        for i, _ in enumerate(model_desmo.sin_coef_list):
            model_desmo.sin_coef_list[i].data = original_sin_coef_list[i].clone()
        for i, _ in enumerate(model_desmo.cos_coef_list):
            model_desmo.cos_coef_list[i].data = original_cos_coef_list[i].clone()
        for i, _ in enumerate(model_desmo.tanh_coef_list):
            model_desmo.tanh_coef_list[i].data = original_tanh_coef_list[i].clone()
        with torch.no_grad():
            model_desmo.c_coef.data[torch.abs(polynorms) < threshold] = 0
            for i, sin_coef in enumerate(model_desmo.sin_coef_list):
                sin_coef.data[torch.abs(nlnorms[i * 3]) < threshold] = 0
            for i, cos_coef in enumerate(model_desmo.cos_coef_list):
                cos_coef.data[torch.abs(nlnorms[i * 3 + 1]) < threshold] = 0
            for i, tanh_coef in enumerate(model_desmo.tanh_coef_list):
                tanh_coef.data[torch.abs(nlnorms[i * 3 + 2]) < threshold] = 0
        model_desmo.eval()
        with torch.no_grad():
            recon, _, _ = model_desmo(snap)
        errs.append(float(np.linalg.norm(X - recon.detach().numpy().T) / np.linalg.norm(X)))
        nonzero = (torch.sum(model_desmo.c_coef != 0).item() + sum(torch.sum(c != 0).item() for c in model_desmo.sin_coef_list) +
                   sum(torch.sum(c != 0).item() for c in model_desmo.cos_coef_list) +
                   sum(torch.sum(c != 0).item() for c in model_desmo.tanh_coef_list))
        counts.append(int(nonzero))
        masks.append(np.concatenate([(model_desmo.c_coef != 0).numpy()] +
                                    [np.array([bool(c != 0) for c in lst]) for lst in
                                     (model_desmo.sin_coef_list, model_desmo.cos_coef_list, model_desmo.tanh_coef_list)]))
    model_desmo.c_coef.data = original_c_coef.clone()
    for lst, orig in ((model_desmo.sin_coef_list, original_sin_coef_list), (model_desmo.cos_coef_list, original_cos_coef_list),
                      (model_desmo.tanh_coef_list, original_tanh_coef_list)):
        for i, _ in enumerate(lst):
            lst[i].data = orig[i].clone()
    return {"sweep_poly_norms": polynorms.numpy().astype(np.float64), "sweep_nl_norms": nlnorms.numpy().astype(np.float64),
            "sweep_thresholds": np.array(THRESHOLDS), "sweep_err": np.array(errs), "sweep_nonzero": np.array(counts),
            "sweep_masks": np.stack(masks)}


ANEU = "DESMO/aneurysm/DESMO_ICA_norm.py"
TURB = "DESMO/turbulent_channel/DESMO-TurbulentChannel.py"


def raw_velocity(n: int, m: int, d: int, seed: int) -> np.ndarray:
    """Reader-shaped input (n*d rows x m snapshots, components of a point on adjacent rows, CYL:52-85); fp32-representable
    values so that the fp32 and fp64 device inputs see the same numbers.  For d = 3 the w rows are NOT zero: the reference
    drops them regardless."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 6.0, m)
    base = rng.standard_normal((n * d, 1)) + 0.6 * rng.standard_normal((n * d, 1)) * np.sin(t)[None, :] \
        + 0.3 * rng.standard_normal((n * d, 1)) * np.cos(2.3 * t)[None, :] + 0.05 * rng.standard_normal((n * d, m))
    return base.astype(np.float32).astype(np.float64)


def preprocess_golden():
    """Runs the reference's own convert3Dto2D_data / convertToMagnitude / subtract_mean (three script variants)."""
    fx = {}
    cyl = ref.load_definitions(ref.CYL, ("convert3Dto2D_data", "convertToMagnitude", "subtract_mean"), {})
    aneu = ref.load_definitions(ANEU, ("convertToMagnitude", "subtract_mean"), {})
    # CYL:170-191 -- 2-D flow stored with 3 components: drop w, magnitude over (u, v), subtract the mean
    X = raw_velocity(301, 37, 3, 11)
    fx["cyl_raw"] = X.copy()
    Y = cyl["convertToMagnitude"](cyl["convert3Dto2D_data"](X.copy()), 2)
    Y, mean = cyl["subtract_mean"](Y)
    fx["cyl_X"], fx["cyl_mean"] = Y, mean
    # ANEU:165-176 -- 3-D flow: magnitude over (u, v, w), subtract the mean, scale by 1/sqrt(m)
    X = raw_velocity(257, 50, 3, 12)
    fx["aneu_raw"] = X.copy()
    Y, mean = aneu["subtract_mean"](aneu["convertToMagnitude"](X.copy()))
    fx["aneu_X"], fx["aneu_mean"] = Y, mean
    # TURB:170-189 -- 3-D magnitude, subtract the mean, then keep every second snapshot
    turb = ref.load_definitions(TURB, ("convertToMagnitude", "subtract_mean"), {})
    X = raw_velocity(192, 41, 3, 13)
    fx["turb_raw"] = X.copy()
    Y, mean = turb["subtract_mean"](turb["convertToMagnitude"](X.copy(), 3))
    fx["turb_X"], fx["turb_mean"] = Y[:, 0::2], mean
    # convertToMagnitude_flag = False (CYL:171): the components stay separate rows, only the mean is removed
    X = raw_velocity(100, 29, 2, 14)
    fx["vec_raw"] = X.copy()
    Y, mean = cyl["subtract_mean"](X.copy())
    fx["vec_X"], fx["vec_mean"] = Y, mean
    fx = {k: (v.astype(np.float32) if k.endswith("_raw") else v) for k, v in fx.items()}  # raw values are fp32-exact by construction
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), **fx)
    print("preprocess.npz", {k: v.shape for k, v in fx.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    preprocess_golden()
    if "--only-preprocess" in sys.argv:
        return
    facts = {"T": {}, "param_totals": {}, "checkpoints": {}}
    for (r, p) in [(4, 3), (4, 2), (2, 2), (8, 3), (8, 2), (32, 2), (3, 4), (2, 7)]:
        ns = ref.load_definitions(ref.CYL, ("binomial_coefficient", "calculate_number_of_terms"), {})
        facts["T"][f"{r},{p}"] = int(ns["calculate_number_of_terms"](r, p))
    # parameter totals printed in the shipped logs (DESMO.out:7-8 of each case)
    facts["param_totals"] = {"CYL": [3961, 1001, 4, 3, None, 62950], "TURB": [16384, 1000, 4, 2, None, 92575],
                             "ANEU": [27000, 1000, 4, 2, None, 135039], "FCYL": [3961, 1001, 2, 2, 10, 8204],
                             "FANEU": [27000, 1000, 2, 2, 30, 54762]}
    for root, _, files in os.walk(ref.REF_ROOT):
        for f in sorted(files):
            if f.endswith(".pt"):
                sd = torch.load(os.path.join(root, f), map_location="cpu", weights_only=True)
                rel = os.path.relpath(os.path.join(root, f), ref.REF_ROOT)
                facts["checkpoints"][rel] = {"keys": list(sd.keys()), "shapes": [list(v.shape) for v in sd.values()],
                                             "numel": int(sum(v.numel() for v in sd.values()))}
    with open(os.path.join(OUT, "facts.json"), "w") as fh:
        json.dump(facts, fh, indent=1)

    # POOL_DATA column order / values straight from the reference
    rng = np.random.default_rng(7)
    pool = {}
    for (r, p) in [(2, 2), (4, 3), (3, 5), (2, 7), (5, 2)]:
        y = rng.standard_normal((9, r)).astype(np.float32)
        ns = ref.load_definitions(ref.CYL, ("POOL_DATA",), {})
        pool[f"y_{r}_{p}"] = y
        pool[f"lib_{r}_{p}"] = ns["POOL_DATA"](torch.from_numpy(y), r, p).numpy()
    np.savez_compressed(os.path.join(OUT, "pool_data.npz"), **pool)

    for (name, kind, n, m, r, p, nF, per0, beta, lam) in GRAD_CASES:
        X, modes = build_inputs(kind, n, m, r)
        snap = np.ascontiguousarray(X.T.astype(np.float32))
        base = orc.init_params(n, m, p, r, nF=nF, period_init=per0 or 60.0)
        prm = orc.perturb(base, seed=42, rel=0.1)
        model, ns = make_model(r, p, n, m, nF, per0, modes)
        load_packed(model, prm)
        mse, ortho, l1, total = ref.reference_losses(model, torch.from_numpy(snap), beta, lam)
        model.zero_grad()
        total.backward()
        g = grads_packed(model, prm)
        with torch.no_grad():
            recon, lat, zv = model(torch.from_numpy(snap))
        fx = {"meta": json.dumps(dict(kind=kind, n=n, m=m, r=r, p=p, nF=nF, period_init=per0, beta=beta, l1_lambda=lam,
                                      data_seed=0, perturb_seed=42, perturb_rel=0.1)),
              "mse": np.float64(mse.item()), "ortho": np.float64(ortho.item()), "l1": np.float64(l1.item()),
              "total": np.float64(total.item()), "recon_sample": recon.numpy()[::7, ::5].copy(),
              "latent": lat.numpy(), "z_values": zv.numpy()}
        for k, v in g.items():
            fx["grad_" + k] = v
        pn, nl = reference_norms(model, ns, nF is not None)   # raw phi_list, as the scripts call them
        fx["poly_norms"] = pn.numpy().astype(np.float64)
        fx["nl_norms"] = nl.numpy().astype(np.float64)  # order: sin_i, cos_i, tanh_i per mode (CYL:686-688)
        np.savez_compressed(os.path.join(OUT, f"grad_{name}.npz"), **fx)
        print(name, "mse", mse.item(), "ortho", ortho.item(), "l1", l1.item())

    for (name, kind, n, m, r, p, nF, per0, beta, lam, steps, om0, lrs, marks) in TRAJ_CASES:
        X, modes = build_inputs(kind, n, m, r)
        snap = torch.from_numpy(np.ascontiguousarray(X.T.astype(np.float32)))
        base = orc.init_params(n, m, p, r, omega_init=om0, nF=nF, period_init=per0 or 60.0)
        prm = orc.perturb(base, seed=43, rel=0.02)  # tiny perturbation: keeps ortho signs away from rounding noise
        model, ns = make_model(r, p, n, m, nF, per0, modes)
        load_packed(model, prm)
        opt = ref.reference_optimizer(model, nF is not None)
        for grp, lr in zip(opt.param_groups, lrs):
            grp["lr"] = lr
        sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", patience=5, factor=0.1, min_lr=1e-6)
        hist = []
        snaps = {}
        for ep in range(steps):
            mse, ortho, l1, total = ref.reference_losses(model, snap, beta, lam)
            opt.zero_grad()
            total.backward()
            opt.step()
            hist.append((mse.item(), ortho.item(), l1.item(), total.item()))
            if ep % 10 == 0:
                sch.step(total.item())  # CYL:776-778 cadence; patience shortened so that LR drops occur inside the fixture
            if ep + 1 in marks:
                sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
                q = orc.from_state_dict(sd, r, p)
                for k in ("gates", "phi", "omega", "zall", "coefs", "periods"):
                    if getattr(q, k) is not None:
                        snaps[f"step{ep + 1}_{k}"] = getattr(q, k)
        fx = {"meta": json.dumps(dict(kind=kind, n=n, m=m, r=r, p=p, nF=nF, period_init=per0, beta=beta, l1_lambda=lam,
                                      data_seed=0, perturb_seed=43, perturb_rel=0.02, steps=steps, patience=5, sched_every=10,
                                      omega_init=om0, lrs=list(lrs), marks=list(marks))),
              "hist": np.array(hist), "final_lrs": np.array([g["lr"] for g in opt.param_groups])}
        fx.update(snaps)
        if steps >= 1000:  # post-hoc sweep on the trained reference module: norms, masks, errors (CYL:1184-1265)
            fx.update(reference_threshold_sweep(model, ns, X, snap, nF is not None))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **fx)
        print(name, "first", hist[0], "last", hist[-1], "lrs", fx["final_lrs"])


if __name__ == "__main__":
    main()
