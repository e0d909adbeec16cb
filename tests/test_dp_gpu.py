"""Data parallel on real GPUs (`-m gpu`, needs >= 2 devices: skipped on the single-GPU test box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`): tests/dp_check.py under torchrun, world_size 2."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def test_two_rank_peer_update_equals_mean_gradient_adam():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29731', os.path.join(ROOT, 'tests', 'dp_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0 and 'DP_CHECK PASS' in r.stdout
