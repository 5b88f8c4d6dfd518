"""Sharded (multi-GPU) parity inside the `-m gpu` run: spawns tests/check_shard_gpu.py on 2 ranks when the box has at
least two GPUs (skipped on a single-GPU box; the driver's scaling run and profiles/r2_shard_parity_n*.log cover
2 / 4 / 8)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("key_bits", ["18", "0"])
def test_two_rank_pipelined_shard_parity(key_bits):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, KEY_BITS=key_bits)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "check_shard_gpu.py")],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "sharded parity OK on 2 GPUs" in r.stdout
