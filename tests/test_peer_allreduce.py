"""Multi-GPU: the one-shot all-reduce over NVLink peer memory (csrc/peer.cu) in the sharded train step -- against the NCCL step, against
the unsharded step, eager and under CUDA-graph replay.  Needs two GPUs on the node (skipped otherwise); one process per GPU (torchrun)."""
import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_allreduce_step_matches_nccl_and_unsharded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one node")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "tests", "_peer_worker.py")]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert pr.returncode == 0, pr.stdout[-3000:] + pr.stderr[-3000:]
    line = [ln for ln in pr.stdout.splitlines() if ln.startswith("PEER_RESULT ")]
    assert line, pr.stdout[-3000:]
    res = json.loads(line[-1][len("PEER_RESULT "):])
    assert res["peer_gates_identical_across_ranks"]
    for k, v in res.items():
        if k.startswith("graph_vs_eager"):
            assert v == 0.0, (k, v)                 # graph replay = the same launches
        elif k.startswith("peer_vs_nccl"):
            assert v < 2e-6, (k, v)                 # only the order of the inter-rank sum differs
        elif k.startswith("peer_vs_unsharded"):
            assert v < 2e-5, (k, v)                 # sharding changes the order of the per-point sums (6 Adamax steps)
