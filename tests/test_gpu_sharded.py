"""Node-sharded online training over 2+ GPUs (BASELINE config 5's layout): one process per GPU, the per-sample min-loc
exchange runs inside the persistent kernel over NVLink peer memory; results must equal the oracle bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import REPO


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_node_sharded_training_matches_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           "29533", os.path.join(REPO, "tests", "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout
