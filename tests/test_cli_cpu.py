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


def test_main_script_equals_reference_main_script(tmp_path):
    """src/main.py run as a script with the reference's default output locations (`../log`, `../model`, `../result`
    relative to the working directory) against the reference's own main.py run the same way
    (tests/golden/main_recmodel.json, oracle/make_golden.py::make_main_fixture): the same files under the same
    hyper-parameter-derived names (src/main.py:63-84), the same log lines (wording, order, %.4f metrics; durations
    masked), the same saved predictions."""
    import json
    import re
    from conftest import GOLDEN
    want = json.load(open(os.path.join(GOLDEN, 'main_recmodel.json')))
    synth.write_dataset(str(tmp_path / 'datasets'), 'toy', 120, 150, 10, feat_dim=64, seed=9)
    os.makedirs(str(tmp_path / 'src'))
    cmd = [sys.executable, os.path.join(ROOT, 'src', 'main.py')] + want['args'] + ['--path', '../datasets/']
    r = subprocess.run(cmd, cwd=str(tmp_path / 'src'), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    files = sorted(os.path.relpath(os.path.join(d, f), str(tmp_path)) for sub in ('log', 'model', 'result')
                   for d, _, fs in os.walk(str(tmp_path / sub)) for f in fs)
    assert files == want['files']
    keep = ('load ', 'size of ', 'label:', '# of ', 'Model # of', 'Drop Neg', 'Prepare ', 'Optimizer:', 'Init:', 'Epoch ',
            'Best Iter', 'Early stop', 'Save model', 'Load model', 'building ', 'loss = ', 'l2 inappropriate', 'Test Before',
            'Test After', 'Save Test Results', '# cuda devices', 'DataLoader:', 'Model:', 'Runner:', 'DataProcessor:')
    log = open(str(tmp_path / [f for f in files if f.startswith('log')][0])).read().split(chr(10))
    log = [re.sub(r'^(INFO|WARNING|ERROR|DEBUG):root:', '', ln).strip() for ln in log]
    ours = [re.sub(r'\[\d+\.\d+ s\]', '[T s]', ln.replace(str(tmp_path), '<root>')) for ln in log if ln.startswith(keep)]
    assert ours == want['log']
    res = np.load(str(tmp_path / [f for f in files if f.startswith('result')][0]))
    assert res.shape == (len(want['result']),)
    assert np.abs(res - np.array(want['result'])).max() <= 1e-6 * np.abs(want['result']).max()
