"""Synthetic datasets in the reference's on-disk formats (SURVEY.md §8d presets).

The reference ships no data; its loader (src/data_loaders/DataLoader.py:79-98) reads headerless
`uid,iid,label,time` CSVs `<path>/<ds>/<ds>.{train,validation,test}.csv`, DCCF reads
`<ds>_<sentence_model>.npy` [I,F] and `<ds>.ips_expo_prob.npy` [U,I] (src/models/DCCF.py:55,64), and the
preprocessing scripts split each user's interactions 70/10/20 (src/data_preprocessing/amazon_data_split_RAND.py:99-122).
"""
import os

import numpy as np

PRESETS = {
    # name: (n_users, n_items, interactions per user)
    'tiny': (2000, 5000, 16),
    'electronics': (48000, 16000, 25),
    'cds': (24000, 20000, 33),
    'yelp': (32000, 38000, 49),
}

DEFAULT_SENTENCE_MODEL = 'paraphrase-distilroberta-base-v1'


def make_interactions(n_users, n_items, per_user, seed=2019, zipf=0.8):
    """Per-user item sets without replacement, item popularity ~ Zipf(zipf); returns dict of
    (uid, iid, label, time) int64 arrays for train / validation / test (70 % / 10 % / 20 % per user)."""
    rs = np.random.RandomState(seed)
    pop = 1.0 / np.arange(1, n_items + 1, dtype=np.float64) ** zipf
    cdf = np.cumsum(pop / pop.sum())
    perm = rs.permutation(n_items)          # popularity rank -> item id
    per_user = min(per_user, n_items)
    draws = np.searchsorted(cdf, rs.random_sample((n_users, per_user * 3)), side='right')
    draws = np.minimum(draws, n_items - 1)
    out = {k: [[], [], [], []] for k in ('train', 'validation', 'test')}
    t = 0
    for u in range(n_users):
        _, first = np.unique(draws[u], return_index=True)
        items = perm[draws[u][np.sort(first)][:per_user]]
        if len(items) < per_user:           # top up with unseen uniform items
            extra = np.setdiff1d(rs.choice(n_items, per_user * 2, replace=False), items)[:per_user - len(items)]
            items = np.concatenate([items, extra])
        n = len(items)
        n_test = max(1, int(round(n * 0.2)))
        n_val = max(1, int(round(n * 0.1)))
        n_train = max(1, n - n_test - n_val)
        labels = rs.randint(1, 6, size=n)
        times = t + np.arange(n)
        t += n
        for name, sl in (('train', slice(0, n_train)), ('validation', slice(n_train, n_train + n_val)),
                         ('test', slice(n_train + n_val, n))):
            o = out[name]
            o[0].append(np.full(len(items[sl]), u, dtype=np.int64))
            o[1].append(items[sl].astype(np.int64))
            o[2].append(labels[sl].astype(np.int64))
            o[3].append(times[sl].astype(np.int64))
    res = {k: tuple(np.concatenate(c) for c in v) for k, v in out.items()}
    return res


def make_features(n_items, feat_dim=768, seed=2019):
    """Roughly unit-norm rows like sentence embeddings."""
    rs = np.random.RandomState(seed + 1)
    return (rs.standard_normal((n_items, feat_dim)) / np.sqrt(feat_dim)).astype(np.float32)


def make_expo_dense(n_users, n_items, seed=2019):
    rs = np.random.RandomState(seed + 2)
    return rs.random_sample((n_users, n_items)).astype(np.float32)


def make_ipsmf_factors(n_users, n_items, dim=64, seed=2019):
    """IPSBiasedMF parameters (src/models/IPSBiasedMF.py:29-35) + propensity, for on-the-fly exposure."""
    rs = np.random.RandomState(seed + 3)
    return {
        'mf_user': (rs.standard_normal((n_users, dim)) * 0.1).astype(np.float32),
        'mf_item': (rs.standard_normal((n_items, dim)) * 0.1).astype(np.float32),
        'mf_user_bias': (rs.standard_normal(n_users) * 0.1).astype(np.float32),
        'mf_item_bias': (rs.standard_normal(n_items) * 0.1).astype(np.float32),
        'mf_global_bias': np.float32(0.1),
        'propensity': rs.random_sample(n_items).astype(np.float32),
        'mf_min_propensity': np.float32(0.1),
    }


def write_dataset(path, dataset, n_users, n_items, per_user, feat_dim=768, seed=2019,
                  sentence_model=DEFAULT_SENTENCE_MODEL, sep=',', ipsmf=False, force_item_max=True):
    """Write the files DataLoader + DCCF read.  Returns the dataset directory."""
    d = os.path.join(path, dataset)
    os.makedirs(d, exist_ok=True)
    inter = make_interactions(n_users, n_items, per_user, seed=seed)
    # item_num = max iid + 1 (DataLoader.py:145-148): make it equal the requested size
    if force_item_max and max(v[1].max() for v in inter.values()) < n_items - 1:
        inter['train'][1][0] = n_items - 1
    for name in ('train', 'validation', 'test'):
        u, i, l, t = inter[name]
        np.savetxt(os.path.join(d, '%s.%s.csv' % (dataset, name)), np.stack([u, i, l, t], axis=1), fmt='%d',
                   delimiter=sep)
    np.save(os.path.join(d, '%s_%s.npy' % (dataset, sentence_model)), make_features(n_items, feat_dim, seed))
    np.save(os.path.join(d, dataset + '.ips_expo_prob.npy'), make_expo_dense(n_users, n_items, seed))
    if ipsmf:
        f = make_ipsmf_factors(n_users, n_items, seed=seed)
        np.savez(os.path.join(d, dataset + '.ipsmf.npz'), **f)
        np.save(os.path.join(d, dataset + '.propensity.npy'), f['propensity'])
    return d


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser(description='write a synthetic DCCF dataset')
    ap.add_argument('--path', default='../datasets/')
    ap.add_argument('--dataset', default='tiny')
    ap.add_argument('--preset', default='tiny', choices=sorted(PRESETS))
    ap.add_argument('--feat_dim', type=int, default=768)
    ap.add_argument('--seed', type=int, default=2019)
    a = ap.parse_args()
    U, I, per = PRESETS[a.preset]
    print(write_dataset(a.path, a.dataset, U, I, per, feat_dim=a.feat_dim, seed=a.seed))
