"""Worker of tests/test_peer_allreduce.py (one process per GPU under torchrun): the sharded train step with the exchange over NVLink peer
memory (csrc/peer.cu) against the same step with NCCL all-reduces and against the unsharded step on one GPU."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desmo_b200 import DesmoEngine, DesmoTrainer  # noqa: E402
from desmo_b200.dist import shard_bounds  # noqa: E402
from tests.helpers import engine_params, load_engine, make_case, rel  # noqa: E402


def slab_engine(prm, modes, snap, lo, hi, n, dev, pg=None):
    import copy

    e = DesmoEngine(hi - lo, prm.m, prm.polyorder, prm.r, device=dev, n_global=n, process_group=pg)
    q = copy.copy(prm)
    q.phi = prm.phi[:, lo:hi]
    q.n = hi - lo
    load_engine(e, q, modes[lo:hi], snap[:, lo:hi])
    return e


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, m, r, p, steps = 5000, 300, 4, 2, 6
    _, modes, snap, prm = make_case("aneurysm", n, m, r, p, omega_init=10.0, perturb_rel=0.05)
    lo, hi = shard_bounds(n, world, rank)
    out = {}
    runs = {}
    for mode in ("nccl", "peer", "peer-graph"):
        e = slab_engine(prm, modes, snap, lo, hi, n, dev, dist.group.WORLD)
        e.set_hyper((1e-2, 1e-3, 1e-2, 1e-2), 1e-3, 1e-4)
        if mode != "nccl":
            assert e.enable_peer_allreduce(), e.peer_status
        if mode == "peer-graph":
            tr = DesmoTrainer(e, lrs=(1e-2, 1e-3, 1e-2, 1e-2), beta=1e-3, l1_lambda=1e-4, sched_every=10 ** 9)
            for _ in range(steps):
                tr.step()
        else:
            for _ in range(steps):
                e.train_step()
        torch.cuda.synchronize()
        runs[mode] = (engine_params(e), e.losses.cpu().numpy().copy(), e.red.cpu().numpy().copy())
        del e
    # replicated parameters must be bit-identical across ranks in peer mode (rank-ordered sums)
    g = torch.from_numpy(runs["peer"][0]["gates"]).to(dev)
    gl = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gl, g)
    out["peer_gates_identical_across_ranks"] = all(torch.equal(gl[0], x) for x in gl)
    for k in ("gates", "zall", "omega", "phi"):
        out[f"peer_vs_nccl_{k}"] = rel(runs["peer"][0][k], runs["nccl"][0][k])
        out[f"graph_vs_eager_{k}"] = rel(runs["peer-graph"][0][k], runs["peer"][0][k])
    out["peer_vs_nccl_red"] = rel(runs["peer"][2], runs["nccl"][2])
    out["peer_vs_nccl_losses"] = rel(runs["peer"][1], runs["nccl"][1])
    if rank == 0:
        e1 = DesmoEngine(n, m, p, r, device=dev)
        load_engine(e1, prm, modes, snap)
        e1.set_hyper((1e-2, 1e-3, 1e-2, 1e-2), 1e-3, 1e-4)
        for _ in range(steps):
            e1.train_step()
        torch.cuda.synchronize()
        ref = engine_params(e1)
        for k in ("gates", "zall", "omega"):
            out[f"peer_vs_unsharded_{k}"] = rel(runs["peer"][0][k], ref[k])
        out["peer_vs_unsharded_phi"] = rel(runs["peer"][0]["phi"], ref["phi"][:, lo:hi])
        out["peer_vs_unsharded_losses"] = rel(runs["peer"][1], e1.losses.cpu().numpy())
        print("PEER_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
