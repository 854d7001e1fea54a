import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return load


def golden_params(g, prefix='init_'):
    p = {k: g[prefix + k] for k in ('E_user', 'E_item', 'W', 'b')}
    p['Feat'] = g['feat']
    p['expo'] = g['expo']
    return p


def rel_err(a, b):
    """max |a-b| / max |b|  — the norm-wise relative error of SURVEY.md Appendix D."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).max()
    return float(np.abs(a - b).max() / (denom if denom > 0 else 1.0))


def config0_problem(root, steps):
    """BASELINE.json configs[0] rebuilt from its seeds alone (tests/golden/config0_train.npz holds only the reference's
    OUTPUTS): dataset (dccf_b200.synth), initial parameters (DCCF + init_paras under the same seed), the first `steps`
    training batches of epoch 0 (the host pipeline, pinned by config0_digest.json).  Returns (g, model, feat, expo, Xs)."""
    import torch
    from dccf_b200 import synth
    from dccf_b200.data_loaders.DataLoader import DataLoader
    from dccf_b200.data_processor.DataProcessor import DataProcessor
    from dccf_b200.models.DCCF import DCCF
    g = np.load(os.path.join(GOLDEN, 'config0_train.npz'), allow_pickle=False)
    seed = int(g['seed'])
    U, I, per = synth.PRESETS['tiny']
    d = synth.write_dataset(root, 'tiny', U, I, per, feat_dim=768, seed=seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    dl = DataLoader(path=root, dataset='tiny', label='label', sep=',')
    model = DCCF(path=d, dataset='tiny', sentence_model=synth.DEFAULT_SENTENCE_MODEL, sample_num=int(g['S']),
                 attribute_num=int(g['A']), std=float(g['std']), label_min=0, label_max=1, feature_num=0,
                 user_num=dl.user_num, item_num=dl.item_num, u_vector_size=64, i_vector_size=64, n_layers=1,
                 random_seed=seed, model_path=os.path.join(root, 'm.pt'))
    model.apply(model.init_paras)
    dl.drop_neg()
    dp = DataProcessor(dl, model, rank=1, test_neg_n=1000)
    dp.get_test_data()
    dp.get_validation_data()
    dp.get_train_data(epoch=-1)
    data = dp.get_train_data(epoch=0)
    batches = dp.prepare_batches(data, int(g['batch_size']), train=True)[:steps]
    Xs = [b['X'].cpu().numpy() for b in batches]
    feat = np.load(os.path.join(d, 'tiny_%s.npy' % synth.DEFAULT_SENTENCE_MODEL))
    expo = np.load(os.path.join(d, 'tiny.ips_expo_prob.npy'))
    return g, model, feat, expo, Xs


def config0_draws(g, n_items, steps):
    """The random inputs of those steps, re-drawn exactly as the reference harness drew them (oracle/ref_harness.py):
    torch CPU generator seeded with seed + 1, per step randint -> normal_ -> bernoulli_.  Yields (sample_item, noise, mask)."""
    import torch
    P = 2 * int(g['batch_size'])
    S, A, std, drop = int(g['S']), int(g['A']), float(g['std']), float(g['dropout'])
    N = P * (S + 1) * A
    torch.manual_seed(int(g['seed']) + 1)
    for _ in range(steps):
        si = torch.randint(n_items, size=(P, S))
        noise = torch.empty((N, 768), dtype=torch.float32).normal_(mean=0.0, std=std)
        mask = torch.empty((N, 64), dtype=torch.float32).bernoulli_(1.0 - drop).div_(1.0 - drop)
        yield si, noise, mask


def config0_eval_problem(root):
    """The evaluation fixture of BASELINE.json configs[0] rebuilt from seeds (tests/golden/config0_eval.npz holds the
    reference's predictions and metrics): the first n_users users of the configs[0] test set, positives + 1000 negatives
    each.  Returns (g, model, feat, expo, data dict of those rows)."""
    import torch
    from dccf_b200 import synth
    from dccf_b200.data_loaders.DataLoader import DataLoader
    from dccf_b200.data_processor.DataProcessor import DataProcessor
    from dccf_b200.models.DCCF import DCCF
    g = np.load(os.path.join(GOLDEN, 'config0_eval.npz'), allow_pickle=False)
    seed = int(g['seed'])
    U, I, per = synth.PRESETS['tiny']
    d = synth.write_dataset(root, 'tiny', U, I, per, feat_dim=768, seed=seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    dl = DataLoader(path=root, dataset='tiny', label='label', sep=',')
    model = DCCF(path=d, dataset='tiny', sentence_model=synth.DEFAULT_SENTENCE_MODEL, sample_num=int(g['S']),
                 attribute_num=int(g['A']), std=float(g['std']), label_min=0, label_max=1, feature_num=0,
                 user_num=dl.user_num, item_num=dl.item_num, u_vector_size=64, i_vector_size=64, n_layers=1,
                 random_seed=seed, model_path=os.path.join(root, 'm.pt'))
    model.apply(model.init_paras)
    dl.drop_neg()
    dp = DataProcessor(dl, model, rank=1, test_neg_n=1000)
    te = dp.get_test_data()
    rows = np.nonzero(np.isin(te['uid'], g['users']))[0]
    sub = {k: np.asarray(te[k])[rows] for k in ('uid', 'iid', 'Y', 'X')}
    sub['sample_id'] = np.arange(len(rows))
    feat = np.load(os.path.join(d, 'tiny_%s.npy' % synth.DEFAULT_SENTENCE_MODEL))
    expo = np.load(os.path.join(d, 'tiny.ips_expo_prob.npy'))
    return g, model, feat, expo, sub


def config0_eval_draws(g, n_items, n_rows):
    """Per evaluation batch (sample_item, noise) as the reference harness drew them: generator seeded with seed + 2, per
    batch randint -> normal_ (dropout is off during evaluation: no mask is drawn)."""
    import torch
    S, A, std, B = int(g['S']), int(g['A']), float(g['std']), int(g['eval_batch_size'])
    torch.manual_seed(int(g['seed']) + 2)
    for a in range(0, n_rows, B):
        P = min(B, n_rows - a)
        si = torch.randint(n_items, size=(P, S))
        noise = torch.empty((P * (S + 1) * A, 768), dtype=torch.float32).normal_(mean=0.0, std=std)
        yield a, a + P, si, noise
