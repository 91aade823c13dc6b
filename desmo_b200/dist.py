"""Point sharding across ranks: each rank owns a contiguous slab of mesh points (rows of U, Phi, P);
W / Z / gates / omega (/ periods) and their Adamax state are replicated; ONE all-reduce per step of the packed fp32 buffer
`red` = [E = G^T R (Kp x mld) | sum r^2 | Phi^T Phi (r x r) | d omega (3r)] (include/desmo_b200.h).  No halo, no all-to-all.

Host-side logic only (usable and tested on CPU with the gloo backend); the device work is in libdesmo_b200.so.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

TILE = 128  # slab boundaries are multiples of the 128-point tile of the fused kernel


def shard_bounds(n: int, world: int, rank: int, tile: int = TILE) -> Tuple[int, int]:
    """[lo, hi) of rank's slab: tiles are dealt out as evenly as possible, earlier ranks take the remainder."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    ntiles = (n + tile - 1) // tile
    base, rem = divmod(ntiles, world)
    t_lo = rank * base + min(rank, rem)
    t_hi = t_lo + base + (1 if rank < rem else 0)
    return min(t_lo * tile, n), min(t_hi * tile, n)


@dataclass(frozen=True)
class RedLayout:
    """Offsets inside the all-reduced buffer (must match desmo_red_count / reduce_partials_kernel)."""
    K: int
    Kp: int
    mld: int
    r: int

    @property
    def e_count(self) -> int:
        return self.Kp * self.mld

    @property
    def loss(self) -> int:
        return self.e_count

    @property
    def gram(self) -> int:
        return self.e_count + 1

    @property
    def domega(self) -> int:
        return self.e_count + 1 + self.r * self.r

    @property
    def count(self) -> int:
        return self.e_count + 1 + self.r * self.r + 3 * self.r

    def pack(self, E: np.ndarray, loss_sum: float, gram: np.ndarray, domega: np.ndarray) -> np.ndarray:
        buf = np.zeros(self.count, np.float32)
        m = E.shape[1]
        buf[:self.e_count].reshape(self.Kp, self.mld)[:self.K, :m] = E
        buf[self.loss] = loss_sum
        buf[self.gram:self.gram + self.r * self.r] = gram.reshape(-1)
        buf[self.domega:self.domega + 3 * self.r] = domega
        return buf

    def unpack(self, buf: np.ndarray, m: int):
        E = buf[:self.e_count].reshape(self.Kp, self.mld)[:self.K, :m]
        return E, float(buf[self.loss]), buf[self.gram:self.gram + self.r * self.r].reshape(self.r, self.r), buf[self.domega:self.domega + 3 * self.r]


def padded_k(K: int) -> int:
    return (K + 15) // 16 * 16


def round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a
