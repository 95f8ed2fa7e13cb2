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


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_run_matches_reference_vectors(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29400 + world), os.path.join(ROOT, "tests", "mp_parity.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MP_PARITY_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
