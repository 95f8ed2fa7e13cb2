"""Two-rank (and more, when the box has them) parity of the partitioned operator with halo exchange over NCCL."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("halo", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_run_matches_reference_vectors(world, halo):
    """halo = p2p: traces stored into the neighbour's buffer by the stage kernel (CUDA IPC peer memory); nccl: pack +
    ncclSend/ncclRecv.  Both must reproduce the reference vectors on every rank's owned dofs."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, DGTD_B200_HALO=halo, DGTD_EXPECT_HALO_MODE="2" if halo == "p2p" else "1",
               DGTD_TEST_PARTITION="metis" if halo == "p2p" else "rcb")      # METIS k-way parts with the fused halo, RCB slabs with NCCL
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world + (10 if halo == "nccl" else 0)), os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "MP_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
