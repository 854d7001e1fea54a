"""The reference command line end to end on a GPU: src/main.py on a small synthetic dataset."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from dccf_b200 import synth

pytestmark = pytest.mark.gpu


def test_main_cli_trains_and_evaluates(tmp_path):
    data_root = str(tmp_path / 'datasets')
    synth.write_dataset(data_root, 'toy', n_users=300, n_items=400, per_user=12, feat_dim=768, seed=1)
    src = os.path.join(ROOT, 'src')
    cmd = [sys.executable, 'main.py', '--rank', '1', '--model_name', 'DCCF', '--optimizer', 'Adam', '--lr', '0.001',
           '--dataset', 'toy', '--path', data_root + '/', '--metric', 'ndcg@5,recall@5,precision@5', '--gpu', '0',
           '--epoch', '2', '--test_neg_n', '100', '--log_file', str(tmp_path / 'log.txt'),
           '--result_file', str(tmp_path / 'result.npy'), '--model_path', str(tmp_path / 'model' / 'm.pt')]
    r = subprocess.run(cmd, cwd=src, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    log = open(str(tmp_path / 'log.txt')).read()
    assert 'Test Before Training' in log and 'Test After Training' in log and 'Epoch     2' in log
    res = np.load(str(tmp_path / 'result.npy'))
    n_test_users = len(set(np.loadtxt(os.path.join(data_root, 'toy', 'toy.test.csv'), delimiter=',')[:, 0].tolist()))
    n_test_rows = len(np.loadtxt(os.path.join(data_root, 'toy', 'toy.test.csv'), delimiter=','))
    assert res.shape == (n_test_rows + 100 * n_test_users,) and np.isfinite(res).all()
    assert os.path.exists(os.path.join(data_root, 'toy', 'rank.csv'))
    assert os.path.exists(str(tmp_path / 'model' / 'm.pt'))
    import torch
    sd = torch.load(str(tmp_path / 'model' / 'm.pt'), map_location='cpu')
    assert sorted(sd.keys()) == ['iid_embeddings.weight', 'mlp.0.bias', 'mlp.0.weight', 'uid_embeddings.weight']
