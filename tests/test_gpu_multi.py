"""Multi-GPU path on real GPUs (skipped with fewer than 2): tools/dp_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_data_parallel_equals_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tools', 'dp_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and 'DP CHECK OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
