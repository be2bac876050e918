"""-m gpu, needs >= 2 GPUs (skipped on the single-GPU test box; run with `gpurun --gpus 2`): ShardedFlatAdam with the fused
reduction + Adam + broadcast kernel (nvls / p2p transports) and the NCCL fallback against torch.optim.Adam on the
rank-averaged gradients -- scripts/check_sharded_adam.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs on one box")
def test_sharded_adam_transports_match_torch_adam():
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "check_sharded_adam.py")],
                       capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:]); sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0
    assert r.stdout.count("PASS") == n
