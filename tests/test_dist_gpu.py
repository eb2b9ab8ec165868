"""Multi-GPU tests of the one exchange step of the path (a15) — need >= 2 GPUs (gpurun --gpus 2); skipped on one GPU."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_dynamic_vfe_train_step_over_nccl():
    """naiveSyncBN1d of the product DynamicVFE in train mode over NCCL + the data-parallel gradient all-reduce against the CPU
    oracle over gloo (tests/dist_vfe_train.py asserts the parity on every rank)."""
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tests', 'dist_vfe_train.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"parity": "ok"' in r.stdout
