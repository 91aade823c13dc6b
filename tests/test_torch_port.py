"""The CPU baseline port (oracle/torch_port.py) computes what the oracle (hence the reference) computes."""
import numpy as np
import pytest

from oracle import desmo_oracle as orc
from tests.helpers import make_case, rel


@pytest.mark.parametrize("nF", [None, 4])
def test_torch_port_matches_oracle(nF):
    torch = pytest.importorskip("torch")
    from oracle.torch_port import TorchPort

    _, modes, snap, prm = make_case("cylinder", 120, 40, 3, 3, nF)
    o = orc.loss_and_grads(prm, modes, snap, 1e-3, 1e-4)
    model = TorchPort(prm, modes)
    mse, ortho, l1, total = model.losses(torch.from_numpy(snap), 1e-3, 1e-4)
    total.backward()
    assert abs(mse.item() - o.mse) < 1e-5 * o.mse and abs(total.item() - o.total) < 1e-5 * o.total
    g = model.packed_grads()
    for k, v in o.grads.items():
        assert rel(g[k], v) < 1e-4, k
