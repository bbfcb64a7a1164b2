"""Multi-PROCESS slab exchange under the driver's pytest: scripts/peer_check.py launched with torchrun, all ranks
sharing cuda:0 (SPHE_ONE_GPU=1: gloo for the plumbing, CUDA IPC for the mailboxes -- the same mapping and the same
device-side flag protocol as one rank per GPU over NVLink).  Bar (printed by the script): positions, velocities,
densities and every terrain row a rank keeps BIT-EQUAL to the single-handle run, sediment conserved exactly."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("world,share", [(2, "window"), (2, "allreduce"), (4, "allreduce")])
def test_peer_check_multiprocess_one_gpu(world, share):
    env = dict(os.environ, SPHE_ONE_GPU="1", TERRAIN_SHARE=share, STEPS="10", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "peer_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_CHECK OK world=%d" % world in r.stdout, r.stdout[-4000:]
