"""Sharded checkpoints (desmo_b200/checkpoint.py): rank-0 gather into the reference key layout, resume at another world size."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from desmo_b200 import checkpoint as ck  # noqa: E402
from desmo_b200.dist import shard_bounds  # noqa: E402


class _Engine:
    """CPU stand-in with the attributes the checkpoint code touches (the real DesmoEngine needs a GPU)."""

    def __init__(self, n, n_global, r, m, lo):
        self.n, self.n_global, self.r, self.m, self.polyorder, self.nF = n, n_global, r, m, 2, 0
        ld = (n + 255) // 256 * 256
        g = torch.arange(lo, lo + n, dtype=torch.float32)
        mk = lambda base: torch.cat([torch.stack([base + 1000.0 * i + g for i in range(r)]), torch.zeros(r, ld - n)], dim=1)  # noqa: E731
        self.phi, self.phi_m, self.phi_u, self.P = mk(0.0), mk(0.25), mk(0.5), mk(0.75)
        self.gates = torch.arange(15 + 3 * r, dtype=torch.float32)


class _Trainer:
    def __init__(self, engine):
        self.engine = engine

    def state_dict(self):
        e = self.engine
        model = {"c_coef": e.gates[:15].clone()}
        model.update({f"phi_list.{i}": e.phi[i, :e.n].clone() for i in range(e.r)})
        model["omega_list.0"] = torch.tensor(3.0)
        return {"model": model, "optimizer": {"phi_m": e.phi_m.clone(), "phi_u": e.phi_u.clone(), "gates_m": e.gates.clone()}, "step": 7,
                "pod_modes": e.P[:, :e.n].clone(), "scheduler": {"lrs": [1e-2], "best": 1.0, "num_bad": 0, "patience": 5}, "epoch": 7,
                "beta": 1e-3, "l1_lambda": 1e-4, "sched_every": 10,
                "shape": {"n": e.n, "m": e.m, "r": e.r, "polyorder": 2, "nF": 0, "n_global": e.n_global}}


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_global, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_global, world, rank)
        sd = ck.gather_trainer_state(_Trainer(_Engine(hi - lo, n_global, 3, 20, lo)))
        out.put((rank, None if sd is None else {"phi": torch.stack([sd["model"][f"phi_list.{i}"] for i in range(3)]).numpy(),
                                                "phi_u": sd["optimizer"]["phi_u"].numpy(), "pod": sd["pod_modes"].numpy(),
                                                "keys": list(sd["model"].keys()), "shape": sd["shape"], "gates_m": sd["optimizer"]["gates_m"].numpy()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_gives_the_full_mesh_state():
    import torch.multiprocessing as mp

    n_global = 128 * 5 + 37
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_global, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1] is None and res[0] is not None
    full = _Engine(n_global, n_global, 3, 20, 0)
    assert np.array_equal(res[0]["phi"], full.phi[:, :n_global].numpy())
    assert np.array_equal(res[0]["phi_u"], full.phi_u[:, :n_global].numpy()) and np.array_equal(res[0]["pod"], full.P[:, :n_global].numpy())
    assert res[0]["keys"] == list(_Trainer(full).state_dict()["model"].keys())  # reference key order preserved
    assert res[0]["shape"]["n"] == n_global and res[0]["shape"]["saved_world_size"] == 2
    assert np.array_equal(res[0]["gates_m"], full.gates.numpy())


def test_scatter_reshards_to_any_world_size():
    n_global = 128 * 7 + 5
    full = ck.gather_trainer_state(_Trainer(_Engine(n_global, n_global, 3, 20, 0)))  # world size 1: the gather is a copy
    for world in (1, 2, 3, 8):
        seen = 0
        for rank in range(world):
            lo, hi = shard_bounds(n_global, world, rank)
            part = ck.scatter_trainer_state(full, rank, world)
            want = _Trainer(_Engine(hi - lo, n_global, 3, 20, lo)).state_dict()
            assert part["shape"]["n"] == hi - lo and part["shape"]["n_global"] == n_global
            for i in range(3):
                assert torch.equal(part["model"][f"phi_list.{i}"], want["model"][f"phi_list.{i}"])
            assert torch.equal(part["optimizer"]["phi_m"], want["optimizer"]["phi_m"][:, :hi - lo])
            assert torch.equal(part["pod_modes"], want["pod_modes"])
            assert torch.equal(part["model"]["c_coef"], want["model"]["c_coef"])
            seen += hi - lo
        assert seen == n_global
    with pytest.raises(ValueError):
        ck.scatter_trainer_state(_Trainer(_Engine(100, 300, 3, 20, 0)).state_dict(), 0, 2)  # a local (slab) state is not a full-mesh file


@pytest.mark.gpu
def test_full_mesh_checkpoint_resumes_exactly_and_reshards(tmp_path):
    """GPU: 30 steps straight == 15 steps, save_checkpoint, fresh objects, load_checkpoint, 15 steps (bit for bit, world size 1);
    and the same file cut into two slabs reproduces the unsharded fused pass (sum of the slabs' `red`, dphi columns)."""
    from desmo_b200 import DESMO, DesmoEngine, DesmoTrainer
    from tests.helpers import engine_params, load_engine, make_case, rel

    _, modes, snap, prm = make_case("channel", 900, 64, 4, 2, omega_init=10.0, perturb_rel=0.02)
    lrs = (1e-2, 1e-3, 1e-2, 1e-2)
    dev = torch.device("cuda:0")

    def fresh():
        model = DESMO(prm.n, prm.m, 2, 4, 10.0, pod_modes=modes, device=dev)
        load_engine(model.engine, prm, modes, snap)
        return model, DesmoTrainer(model, lrs=lrs, patience=2, sched_every=5, use_cuda_graph=False)

    m1, t1 = fresh()
    for _ in range(30):
        t1.step()
    m2, t2 = fresh()
    for _ in range(15):
        t2.step()
    path = os.path.join(tmp_path, "full.pt")
    ck.save_checkpoint(t2, path)
    sd = torch.load(path, weights_only=False)
    assert list(sd["model"].keys()) == list(m1.state_dict().keys()) and sd["model"]["phi_list.0"].shape == (prm.n,)
    m3, t3 = fresh()
    m3.engine.P.zero_()
    ck.load_checkpoint(t3, path, map_location=dev)
    for _ in range(15):
        t3.step()
    torch.cuda.synchronize()
    for k, v in engine_params(m1.engine).items():
        assert np.array_equal(v, engine_params(m3.engine)[k]), k
    # re-shard the file into two slabs: the slabs' partial `red` sum to the full pass, dphi columns match
    m2.engine.build_w(False)
    m2.engine.fused_residual_grad()
    reds, dphis = [], []
    for rank in range(2):
        lo, hi = shard_bounds(prm.n, 2, rank)
        part = ck.scatter_trainer_state(sd, rank, 2)
        e = DesmoEngine(hi - lo, prm.m, 2, 4, omega_init=10.0, device=dev, n_global=prm.n)
        tr = DesmoTrainer(e, lrs=lrs, patience=2, sched_every=5, use_cuda_graph=False)
        part["model"] = None  # bare engine: parameters come through the packed buffers below
        tr.load_state_dict(part)
        e.phi.zero_()
        e.phi[:, :hi - lo] = torch.stack([sd["model"][f"phi_list.{i}"][lo:hi] for i in range(4)]).to(dev)
        e.gates.copy_(m2.engine.gates); e.rows.copy_(m2.engine.rows); e.omega.copy_(m2.engine.omega)
        e.set_snapshot(torch.from_numpy(snap[:, lo:hi].copy()))
        assert torch.equal(e.P[:, :hi - lo], m2.engine.P[:, lo:hi]) and torch.equal(e.phi_u[:, :hi - lo], m2.engine.phi_u[:, lo:hi])
        e.build_w(False)
        e.fused_residual_grad()
        reds.append(e.red.clone())
        dphis.append(e.dphi[:, :hi - lo].clone())
    assert rel((reds[0] + reds[1]).cpu().numpy(), m2.engine.red.cpu().numpy()) < 2e-6
    assert rel(torch.cat(dphis, dim=1).cpu().numpy(), m2.engine.dphi[:, :prm.n].cpu().numpy()) < 2e-6
