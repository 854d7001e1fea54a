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


def test_cli_data_parallel_training(tmp_path):
    """src/main.py under torchrun on 2 GPUs: trains, evaluates (user-sharded), rank 0 writes the files, and the
    run ends with the replica checksum (identical parameters on both ranks)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    from dccf_b200 import synth
    data_root = str(tmp_path / 'datasets')
    synth.write_dataset(data_root, 'toy', n_users=300, n_items=400, per_user=12, feat_dim=768, seed=1)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29519', 'main.py', '--rank', '1', '--model_name', 'DCCF', '--optimizer', 'Adam',
           '--lr', '0.001', '--dataset', 'toy', '--path', data_root + '/', '--metric', 'ndcg@5,recall@5,precision@5',
           '--epoch', '2', '--batch_size', '32', '--test_neg_n', '100', '--log_file', str(tmp_path / 'log.txt'),
           '--result_file', str(tmp_path / 'result.npy'), '--model_path', str(tmp_path / 'model' / 'm.pt')]
    r = subprocess.run(cmd, cwd=os.path.join(ROOT, 'src'), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(str(tmp_path / 'log.txt')).read()
    assert 'data parallel: 2 ranks' in log and 'parameter checksums identical on all 2 ranks' in log
    assert 'Test After Training' in log and 'Epoch     2' in log
    assert os.path.exists(str(tmp_path / 'result.npy')) and os.path.exists(str(tmp_path / 'model' / 'm.pt'))
