"""Multi-GPU parity under pytest: launches tests/dist_check.py with torchrun on every power-of-two rank count the box offers
(2, 4, 8), once with the NCCL ghost exchange and once with the peer-memory exchange (DKT_DIST_P2P=1).  The script gathers
the partitioned results to the single-rank node order and compares them with the golden vectors taken from the reference.
Skipped on a box with one GPU."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("p2p", ["0", "1"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_partitioned_matvec_matches_reference(dkt, nranks, p2p):
    if _ngpu() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    env = dict(os.environ)
    env["DKT_DIST_P2P"] = p2p
    if p2p == "1":
        env["DKT_P2P_CHECK"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
