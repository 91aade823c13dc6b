"""N > 1 path on CPU (gloo, world_size 2): point sharding + the single packed all-reduce reproduce the unsharded step.

Each rank evaluates ITS slab with the CPU oracle (standing in for the CUDA kernels, which need a GPU), packs the partials exactly
as libdesmo_b200 lays out `red`, all-reduces, and applies the replicated update; the result must equal the single-process oracle."""
import os
import socket

import numpy as np
import pytest

from desmo_b200.dist import RedLayout, padded_k, round_up, shard_bounds
from oracle import desmo_oracle as orc
from tests.helpers import make_case, rel


def test_shard_bounds_cover_and_align():
    for n in (1, 127, 128, 129, 1000, 3961, 27000, 3 * 2 ** 20):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for rank in range(world):
                lo, hi = shard_bounds(n, world, rank)
                assert lo == prev and lo <= hi <= n
                assert lo % 128 == 0 or lo == n
                prev = hi
                sizes.append(hi - lo)
            assert prev == n
            assert max(sizes) - min(sizes) < 256 or n < 128 * world  # at most one tile plus the ragged tail


def test_red_layout_matches_library():
    import ctypes

    from desmo_b200 import _lib

    lib = _lib.load()
    for (n, m, r, p) in [(1000, 100, 4, 2), (3961, 1001, 4, 3), (500, 64, 2, 2)]:
        K = orc.number_of_terms(r, p) + 3 * r
        lay = RedLayout(K, padded_k(K), round_up(m, 16), r)
        assert lay.count == lib.desmo_red_count(ctypes.byref(_lib.make_shape(n, m, r, p)))
        assert lay.Kp == lib.desmo_padded_k(r, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _slab_partials(prm, modes, snap, lo, hi, n_global):
    """What one rank's fused pass produces for points [lo, hi): E, sum r^2, Phi^T Phi, d omega (scaled by 2/(n_global m)), d phi."""
    q = prm.copy()
    q.n, q.phi = hi - lo, prm.phi[:, lo:hi].copy()
    G, lat = orc.spatial_library(q, modes[lo:hi])
    W = (q.gates[:, None] * orc.temporal_rows(q)).astype(np.float32)
    R = (G @ W - snap[:, lo:hi].T).astype(np.float32)
    scale = np.float32(2.0 / (n_global * prm.m))
    D = scale * (R @ W.T)
    T, r = q.T, q.r
    om = q.omega
    cs, sn, th = np.cos(om[0::3] * lat), np.sin(om[1::3] * lat), np.tanh(om[2::3] * lat)
    sech2 = 1.0 - th * th
    Ds, Dc, Dh = D[:, T:T + r], D[:, T + r:T + 2 * r], D[:, T + 2 * r:]
    dlat = orc.pool_data_derivative(lat, q.polyorder, D[:, :T]) + Ds * om[0::3] * cs - Dc * om[1::3] * sn + Dh * om[2::3] * sech2
    domega = np.zeros(3 * r, np.float32)
    domega[0::3], domega[1::3], domega[2::3] = (Ds * lat * cs).sum(0), -(Dc * lat * sn).sum(0), (Dh * lat * sech2).sum(0)
    return G.T @ R, float((R.astype(np.float64) ** 2).sum()), lat.T @ lat, domega, dlat, lat


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, modes, snap, prm = make_case("channel", 700, 48, 4, 2)
        beta, lam = 1e-3, 1e-4
        lay = RedLayout(prm.K, padded_k(prm.K), round_up(prm.m, 16), prm.r)
        lo, hi = shard_bounds(prm.n, world, rank)
        E, loss, gram, domega, dlat, lat = _slab_partials(prm, modes, snap, lo, hi, prm.n)
        red = torch.from_numpy(lay.pack(E, loss, gram, domega))
        dist.all_reduce(red)  # the ONE collective of the step
        E, loss, gram, domega = lay.unpack(red.numpy(), prm.m)
        # replicated part of the update (identical on every rank) + the rank's own phi slab
        scale = np.float32(2.0 / (prm.n * prm.m))
        zrows = orc.temporal_rows(prm)
        grads = {"zall": prm.gates[:, None] * (scale * E), "gates": (zrows * (scale * E)).sum(1) + lam * np.sign(prm.gates), "omega": domega}
        sgn = np.sign(gram)
        np.fill_diagonal(sgn, 0.0)
        dphi = ((dlat + np.float32(beta) * (lat @ sgn.T)) * modes[lo:hi, :prm.r]).T.astype(np.float32)
        ref = orc.loss_and_grads(prm, modes, snap, beta, lam)
        ok = (abs(loss / (prm.n * prm.m) - ref.mse) < 1e-5 * ref.mse and rel(grads["zall"], ref.grads["zall"]) < 1e-5 and
              rel(grads["gates"], ref.grads["gates"]) < 1e-5 and rel(grads["omega"], ref.grads["omega"]) < 5e-5 and
              rel(dphi, ref.grads["phi"][:, lo:hi]) < 5e-5)
        # every rank must hold bit-identical replicated gradients after the all-reduce
        gathered = [torch.zeros_like(red) for _ in range(world)]
        dist.all_gather(gathered, red)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        out.put((rank, bool(ok), bool(same)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_step_equals_single_process():
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] for r in res), res
