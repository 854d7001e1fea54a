"""The reference's command line on CPU for the models that do not need the CUDA kernels (RecModel: plain matrix
factorisation through the autograd path of the mirrored runner): loader -> processor -> fit -> evaluate -> model
selection -> checkpoint -> result file, i.e. everything of src/main.py's flow except DCCF's device math."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from dccf_b200 import synth


@pytest.mark.parametrize('rank,metric', [(1, 'rmse'), (0, 'rmse,mae')])
def test_main_cli_runs_recmodel_on_cpu(tmp_path, rank, metric):
    data_root = str(tmp_path / 'datasets')
    synth.write_dataset(data_root, 'toy', n_users=120, n_items=150, per_user=10, feat_dim=64, seed=2)
    cmd = [sys.executable, 'main.py', '--rank', str(rank), '--model_name', 'RecModel', '--optimizer', 'Adam', '--lr', '0.01',
           '--dataset', 'toy', '--path', data_root + '/', '--metric', metric, '--gpu', '', '--epoch', '2',
           '--batch_size', '64', '--test_neg_n', '5', '--log_file', str(tmp_path / 'log.txt'),
           '--result_file', str(tmp_path / 'result.npy'), '--model_path', str(tmp_path / 'model' / 'm.pt')]
    r = subprocess.run(cmd, cwd=os.path.join(ROOT, 'src'), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = open(str(tmp_path / 'log.txt')).read()
    for needle in ('Test Before Training', 'Epoch     1', 'Epoch     2', 'Best Iter(validation)', 'Test After Training',
                   'Save Test Results'):
        assert needle in log, needle
    res = np.load(str(tmp_path / 'result.npy'))
    n_test = len(np.loadtxt(os.path.join(data_root, 'toy', 'toy.test.csv'), delimiter=','))
    n_users = len(set(np.loadtxt(os.path.join(data_root, 'toy', 'toy.test.csv'), delimiter=',')[:, 0].tolist()))
    assert res.shape == ((n_test + 5 * n_users,) if rank == 1 else (n_test,)) and np.isfinite(res).all()
    assert os.path.exists(str(tmp_path / 'model' / 'm.pt'))
    assert os.path.exists(os.path.join(data_root, 'toy', 'rank.csv'))
    for generated in ('toy.info.json', 'toy.train_group.csv', 'toy.vt_group.csv'):
        assert os.path.exists(os.path.join(data_root, 'toy', generated))
