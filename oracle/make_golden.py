"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src, through oracle/ref_harness.py) on CPU.  Run in the build container only:

    python oracle/make_golden.py

The fixtures are committed; the GPU box never needs the reference.
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from dccf_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SENT = synth.DEFAULT_SENTENCE_MODEL


def _params_of(model):
    return {
        'E_user': model.uid_embeddings.weight.detach().numpy().copy(),
        'E_item': model.iid_embeddings.weight.detach().numpy().copy(),
        'W': model.mlp[0].weight.detach().numpy().copy(),
        'b': model.mlp[0].bias.detach().numpy().copy(),
    }


class _StubDP:
    """prepare_batches() provider for BaseRunner.fit / predict (BaseRunner.py:170,143)."""
    rank = 1

    def __init__(self, batches):
        self.batches = batches

    def prepare_batches(self, data, batch_size, train):
        return self.batches


def make_train_fixture(name, U, I, F, P, steps, S=10, A=2, std=0.1, dropout=0.2, seed=2019, l2=1e-4, lr=1e-3):
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        d = os.path.join(tmp, 'g')
        os.makedirs(d)
        rs = np.random.RandomState(seed)
        feat = (rs.standard_normal((I, F)) / np.sqrt(F)).astype(np.float32)
        expo = rs.random_sample((U, I)).astype(np.float32)
        np.save(os.path.join(d, 'g_%s.npy' % SENT), feat)
        np.save(os.path.join(d, 'g.ips_expo_prob.npy'), expo)
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            model = rh.build_reference_model(ref, d, 'g', SENT, U, I, sample_num=S, attribute_num=A, std=std,
                                             random_seed=seed, model_path=os.path.join(tmp, 'm.pt'))
            init = _params_of(model)
            b = P // 2
            batches, Xs = [], []
            for t in range(steps):
                u = rs.randint(0, U, size=b)
                pos = rs.randint(0, I, size=b)
                neg = rs.randint(0, I, size=b)
                X = np.concatenate([np.stack([u, pos], 1), np.stack([u, neg], 1)]).astype(np.int64)
                Y = np.concatenate([np.ones(b, np.float32), np.zeros(b, np.float32)])
                Xs.append(X)
                batches.append({'X': torch.from_numpy(X), 'Y': torch.from_numpy(Y), 'rank': 1,
                                'sample_id': np.arange(2 * b), 'real_batch_size': b, 'total_batch_size': 2 * b})
            outs = []
            model.register_forward_hook(lambda m, i, o: outs.append(
                (o['prediction'].detach().numpy().copy(), float(o['loss'].detach()))))
            runner = ref.BaseRunner(optimizer='Adam', learning_rate=lr, epoch=1, batch_size=b, eval_batch_size=16384,
                                    dropout=dropout, l2=l2, metrics='ndcg@5', check_epoch=1, early_stop=1)
            rh.reset_tape()
            runner.fit(model, None, _StubDP(batches), epoch=0)
            calls = rh.tape().calls
            final = _params_of(model)
            opt = model.optimizer
            names = ['E_user', 'E_item', 'W', 'b']
            plist = [model.uid_embeddings.weight, model.iid_embeddings.weight, model.mlp[0].weight, model.mlp[0].bias]
            out = {'feat': feat, 'expo': expo, 'S': S, 'A': A, 'std': std, 'dropout': dropout, 'l2': l2, 'lr': lr,
                   'steps': steps, 'seed': seed}
            for k, v in init.items():
                out['init_' + k] = v
            for k, v in final.items():
                out['final_' + k] = v
            for n, p in zip(names, plist):
                out['m_' + n] = opt.state[p]['exp_avg'].numpy().copy()
                out['v_' + n] = opt.state[p]['exp_avg_sq'].numpy().copy()
                out['lastgrad_' + n] = p.grad.numpy().copy()    # after l2 term and clip (BaseRunner.py:181-185)
            for t in range(steps):
                out['X_%d' % t] = Xs[t]
                out['sample_item_%d' % t] = calls[t]['sample_item'].numpy()
                out['noise_%d' % t] = calls[t]['noise'].numpy()
                if dropout > 0:
                    out['mask_%d' % t] = calls[t]['masks'][0].numpy()
                out['pred_%d' % t] = outs[t][0]
                out['loss_%d' % t] = outs[t][1]

            # an eval-mode predict through BaseRunner.predict (dropout off, ragged batch)
            Pe = 37
            Xe = np.stack([rs.randint(0, U, size=Pe), rs.randint(0, I, size=Pe)], 1).astype(np.int64)
            ebatch = {'X': torch.from_numpy(Xe), 'Y': torch.zeros(Pe), 'rank': 1, 'sample_id': np.arange(Pe)}
            rh.reset_tape()
            pe = runner.predict(model, {'sample_id': np.arange(Pe)}, _StubDP([ebatch]))
            c = rh.tape().calls[0]
            out['eval_X'] = Xe
            out['eval_sample_item'] = c['sample_item'].numpy()
            out['eval_noise'] = c['noise'].numpy()
            out['eval_pred'] = np.asarray(pe, dtype=np.float32)
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, 'losses', [outs[t][1] for t in range(steps)])
    finally:
        shutil.rmtree(tmp)


def make_sampler_fixture(name='sampler', U=30, I=40, seed=2019, test_neg_n=5, epochs=2, batch_size=16):
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        d = os.path.join(tmp, 's')
        os.makedirs(d)
        rs = np.random.RandomState(seed + 7)
        rows = {'train': [], 'validation': [], 'test': []}
        t = 0
        for u in range(U):
            # user 0: 17 train rows of 40 items -> the per-epoch negative pool shrinks below 20 % (np.random.choice
            # branch, DP:490-493,505-512) during TRAIN sampling; user 1: 34 items overall -> the same branch when
            # the TEST negatives are drawn (DP:482)
            if u == 0:
                n_train, n_val, n_test = 17, 1, 2
            elif u == 1:
                n_train, n_val, n_test = 14, 10, 10
            else:
                n_train, n_val, n_test = int(rs.randint(3, 9)), 1, (2 if u % 3 else 1)
            n = n_train + n_val + n_test
            items = rs.choice(I, n, replace=False)
            for j, it in enumerate(items):
                split = 'train' if j < n_train else ('validation' if j < n_train + n_val else 'test')
                rows[split].append((u, int(it), int(rs.randint(1, 6)), t))
                t += 1
        for k, v in rows.items():
            np.savetxt(os.path.join(d, 's.%s.csv' % k), np.array(v, dtype=np.int64), fmt='%d', delimiter=',')
        np.save(os.path.join(d, 's_%s.npy' % SENT), np.zeros((I, 64), np.float32))
        np.save(os.path.join(d, 's.ips_expo_prob.npy'), np.zeros((U, I), np.float32))
        out = {'train_csv': np.array(rows['train']), 'validation_csv': np.array(rows['validation']),
               'test_csv': np.array(rows['test']), 'seed': seed, 'test_neg_n': test_neg_n, 'epochs': epochs,
               'batch_size': batch_size}
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            dl = ref.DataLoader(path=tmp, dataset='s', label='label', sep=',')
            model = rh.build_reference_model(ref, d, 's', SENT, dl.user_num, dl.item_num, random_seed=seed,
                                             model_path=os.path.join(tmp, 'm.pt'))
            dl.drop_neg()
            dp = ref.DataProcessor(dl, model, rank=1, test_neg_n=test_neg_n)
            out['user_num'], out['item_num'] = dl.user_num, dl.item_num
            te = dp.get_test_data()                    # main.py:181-182 order: test first,
            va = dp.get_validation_data()              # then validation (BaseRunner.py:223)
            for nm, dd in (('test', te), ('validation', va)):
                for k in ('uid', 'iid', 'Y', 'X', 'sample_id'):
                    out['%s_%s' % (nm, k)] = np.asarray(dd[k])
            tr = dp.get_train_data(epoch=-1)
            out['train0_X'] = np.asarray(tr['X']).copy()
            for ep in range(epochs):
                data = dp.get_train_data(epoch=ep)
                out['train_ep%d_order' % ep] = np.asarray(data['sample_id']).copy()
                batches = dp.prepare_batches(data, batch_size, train=True)
                out['train_ep%d_X' % ep] = np.concatenate([b['X'].numpy() for b in batches])
                out['train_ep%d_Y' % ep] = np.concatenate([b['Y'].numpy() for b in batches])
                out['train_ep%d_sample_id' % ep] = np.concatenate([b['sample_id'] for b in batches])
                out['train_ep%d_nbatches' % ep] = len(batches)
            out['np_state_after'] = np.random.get_state()[1][:8].astype(np.int64)
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, 'test rows', len(out['test_X']), 'train rows', len(out['train0_X']))
    finally:
        shutil.rmtree(tmp)


def array_digest(a):
    """sha256 over dtype, shape and bytes of an array (what tests/golden/config0_digest.json stores)."""
    import hashlib
    a = np.ascontiguousarray(a)
    h = hashlib.sha256()
    h.update(str(a.dtype).encode() + b'|' + str(a.shape).encode() + b'|')
    h.update(a.tobytes())
    return h.hexdigest()


def make_config0_digest(name='config0_digest', seed=2019, epochs=2, test_neg_n=1000, batch_size=128):
    """BASELINE.json configs[0] — the reference's own CPU-runnable case: synthetic Electronics-shaped data with 2 000
    users x 5 000 items x 768-d features, test_neg_n = 1000 — through the UNMODIFIED reference's DataLoader /
    DataProcessor (src/data_loaders/DataLoader.py, src/data_processor/DataProcessor.py).  The arrays are too large to
    commit (2 M negatives per evaluation set), so the fixture holds their sha256 digests: every id the host side of
    the path produces at this scale must match bit for bit (tests/test_host_parity.py::test_config0_host_pipeline_digest)."""
    import json
    from dccf_b200 import synth
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        U, I, per = synth.PRESETS['tiny']
        d = synth.write_dataset(tmp, 'tiny', U, I, per, feat_dim=768, seed=seed)
        out = {'preset': 'tiny', 'users': U, 'items': I, 'per_user': per, 'feat_dim': 768, 'seed': seed, 'epochs': epochs,
               'test_neg_n': test_neg_n, 'batch_size': batch_size, 'digests': {}, 'rows': {}}
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            dl = ref.DataLoader(path=tmp, dataset='tiny', label='label', sep=',')
            model = rh.build_reference_model(ref, d, 'tiny', SENT, dl.user_num, dl.item_num, random_seed=seed,
                                             model_path=os.path.join(tmp, 'm.pt'))
            dl.drop_neg()
            dp = ref.DataProcessor(dl, model, rank=1, test_neg_n=test_neg_n)
            out['user_num'], out['item_num'] = int(dl.user_num), int(dl.item_num)
            te = dp.get_test_data()
            va = dp.get_validation_data()
            for nm, dd in (('test', te), ('validation', va)):
                out['rows'][nm] = int(len(dd['Y']))
                for k in ('uid', 'iid', 'Y', 'X', 'sample_id'):
                    out['digests']['%s_%s' % (nm, k)] = array_digest(np.asarray(dd[k]))
            dp.get_train_data(epoch=-1)
            for ep in range(epochs):
                data = dp.get_train_data(epoch=ep)
                batches = dp.prepare_batches(data, batch_size, train=True)
                out['rows']['train_ep%d_batches' % ep] = len(batches)
                out['digests']['train_ep%d_X' % ep] = array_digest(np.concatenate([b['X'].numpy() for b in batches]))
                out['digests']['train_ep%d_Y' % ep] = array_digest(np.concatenate([b['Y'].numpy() for b in batches]))
                out['digests']['train_ep%d_sample_id' % ep] = array_digest(
                    np.concatenate([np.asarray(b['sample_id']) for b in batches]).astype(np.int64))
            out['np_state_after'] = [int(x) for x in np.random.get_state()[1][:8]]
        # the files a first run of DataLoader generates next to the dataset (DataLoader.py:113-129, 169-177): bytes
        import hashlib
        out['files'] = {f: hashlib.sha256(open(os.path.join(d, f), 'rb').read()).hexdigest()
                        for f in ('tiny.info.json', 'tiny.train_group.csv', 'tiny.vt_group.csv')}
        with open(os.path.join(GOLDEN, name + '.json'), 'w') as f:
            json.dump(out, f, indent=1, sort_keys=True)
        print(name, out['rows'])
    finally:
        shutil.rmtree(tmp)


def make_config0_train(name='config0_train', seed=2019, steps=24, batch_size=128, lr=1e-3, l2=1e-4, dropout=0.2):
    """BASELINE.json configs[0] again, now the device math: the first `steps` training steps of epoch 0 of the UNMODIFIED
    reference (its DataLoader -> DataProcessor -> DCCF -> BaseRunner.fit with Adam) at full size — 256 pairs = 5 632
    predictor rows x 832 inputs per step.  The random inputs are NOT stored (17 MB of noise per step): the torch CPU
    generator is re-seeded with seed + 1 right before fit, and a test re-draws them with the same three calls per step
    (torch.randint, normal_, bernoulli_ — oracle/ref_harness.py) on the same torch build.  Stored: every step's loss,
    the predictions of the first and last step, the final bias, the final rows of the users / items of the last batch,
    a slice of W, and the norms of all four final tensors."""
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        U, I, per = synth.PRESETS['tiny']
        d = synth.write_dataset(tmp, 'tiny', U, I, per, feat_dim=768, seed=seed)
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            dl = ref.DataLoader(path=tmp, dataset='tiny', label='label', sep=',')
            model = rh.build_reference_model(ref, d, 'tiny', SENT, dl.user_num, dl.item_num, random_seed=seed,
                                             model_path=os.path.join(tmp, 'm.pt'))
            dl.drop_neg()
            dp = ref.DataProcessor(dl, model, rank=1, test_neg_n=1000)
            dp.get_test_data()                       # the numpy stream is consumed in main.py's order
            dp.get_validation_data()
            dp.get_train_data(epoch=-1)
            data = dp.get_train_data(epoch=0)
            batches = dp.prepare_batches(data, batch_size, train=True)[:steps]
            outs = []
            model.register_forward_hook(lambda m, i, o: outs.append(
                (o['prediction'].detach().numpy().copy(), float(o['loss'].detach()))))
            runner = ref.BaseRunner(optimizer='Adam', learning_rate=lr, epoch=1, batch_size=batch_size,
                                    eval_batch_size=16384, dropout=dropout, l2=l2, metrics='ndcg@5', check_epoch=1,
                                    early_stop=1)
            rh.reset_tape()
            torch.manual_seed(seed + 1)              # the replay point of the test
            runner.fit(model, None, _StubDP(batches), epoch=0)
            final = _params_of(model)
            X_last = batches[-1]['X'].numpy()
            users, items = np.unique(X_last[:, 0]), np.unique(X_last[:, 1])
            out = {'seed': seed, 'steps': steps, 'batch_size': batch_size, 'lr': lr, 'l2': l2, 'dropout': dropout,
                   'std': 0.1, 'S': 10, 'A': 2, 'torch_version': np.array(torch.__version__),
                   'loss': np.array([o[1] for o in outs], dtype=np.float64), 'pred_first': outs[0][0],
                   'pred_last': outs[-1][0], 'X_first': batches[0]['X'].numpy(), 'X_last': X_last,
                   'sample_item_first': rh.tape().calls[0]['sample_item'].numpy(),
                   'final_b': final['b'], 'users_last': users, 'items_last': items,
                   'final_E_user_rows': final['E_user'][users], 'final_E_item_rows': final['E_item'][items],
                   'final_W_rows': final['W'][::8], 'norms': np.array([np.linalg.norm(final[k].astype(np.float64))
                                                                       for k in ('E_user', 'E_item', 'W', 'b')])}
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, 'losses', out['loss'][:3], '...', out['loss'][-1])
    finally:
        shutil.rmtree(tmp)


def make_config0_eval(name='config0_eval', seed=2019, n_users=100, eval_batch_size=16384):
    """BASELINE.json configs[0], the evaluation half: the first `n_users` test users of the configs[0] test set (each
    with its positives + 1000 sampled negatives) scored by the UNMODIFIED reference (BaseRunner.predict ->
    DCCF.predict, batches of 16 384 pairs = 360 448 predictor rows) and ranked by its evaluate_method
    (ndcg@5 / recall@5 / precision@5).  As in make_config0_train the random inputs are re-drawable (torch CPU
    generator seeded with seed + 2 right before predict: per batch randint -> normal_), only outputs are stored."""
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        U, I, per = synth.PRESETS['tiny']
        d = synth.write_dataset(tmp, 'tiny', U, I, per, feat_dim=768, seed=seed)
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            dl = ref.DataLoader(path=tmp, dataset='tiny', label='label', sep=',')
            model = rh.build_reference_model(ref, d, 'tiny', SENT, dl.user_num, dl.item_num, random_seed=seed,
                                             model_path=os.path.join(tmp, 'm.pt'))
            dl.drop_neg()
            dp = ref.DataProcessor(dl, model, rank=1, test_neg_n=1000)
            te = dp.get_test_data()
            uid = np.asarray(te['uid'])
            _, first = np.unique(uid, return_index=True)
            users = uid[np.sort(first)][:n_users]
            rows = np.nonzero(np.isin(uid, users))[0]
            sub = {k: np.asarray(te[k])[rows] for k in ('uid', 'iid', 'Y', 'X')}
            sub['sample_id'] = np.arange(len(rows))
            runner = ref.BaseRunner(optimizer='Adam', learning_rate=1e-3, epoch=1, batch_size=128,
                                    eval_batch_size=eval_batch_size, dropout=0.2, l2=1e-4,
                                    metrics='ndcg@5,recall@5,precision@5', check_epoch=1, early_stop=1)
            rh.reset_tape()
            torch.manual_seed(seed + 2)              # the replay point of the test
            pred = runner.predict(model, sub, dp)
            metrics = ['ndcg@5', 'recall@5', 'precision@5']
            vals = model.evaluate_method(pred, sub, metrics=metrics)
            out = {'seed': seed, 'n_users': n_users, 'eval_batch_size': eval_batch_size, 'users': users,
                   'n_rows': len(rows),
                   'torch_version': np.array(torch.__version__), 'std': 0.1, 'S': 10, 'A': 2,
                   'X_digest': np.array(array_digest(sub['X'])), 'pred': np.asarray(pred, dtype=np.float32),
                   'sample_item_first_digest': np.array(array_digest(rh.tape().calls[0]['sample_item'].numpy())),
                   'metrics': np.array(metrics), 'values': np.array([float(v) for v in vals], dtype=np.float64)}
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, len(rows), 'rows', dict(zip(metrics, vals)))
    finally:
        shutil.rmtree(tmp)


def make_ipsmf_fixture(name='ipsmf', U=70, I=110, seed=2019, M=0.1):
    """IPSBiasedMF.predict of the UNMODIFIED reference (src/models/IPSBiasedMF.py:37-57) over the whole U x I grid:
    the exposure source of the scaled configuration and the producer of ips_expo_prob.npy (README.md:27-29)."""
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        d = os.path.join(tmp, 'p')
        os.makedirs(d)
        rs = np.random.RandomState(seed + 3)
        prop = rs.random_sample(I).astype(np.float32)
        prop[:5] = [0.0, 0.05, 0.1, 0.1000001, 0.5]            # below, at and above the clamp M
        np.save(os.path.join(d, 'p' + ref.global_p.PROPENSITY_SUFFIX), prop)
        with rh.cpu_shims():
            torch.manual_seed(seed)
            model = ref.IPSBiasedMF(path=d, dataset='p', M=M, label_min=0, label_max=1, feature_num=0, user_num=U,
                                    item_num=I, u_vector_size=64, i_vector_size=64, random_seed=seed,
                                    model_path=os.path.join(tmp, 'm.pt'))
            with torch.no_grad():                              # biases away from their init so that every term counts
                for p_ in model.parameters():
                    p_.copy_(torch.from_numpy(np.asarray(rs.standard_normal(tuple(p_.shape)) * 0.3, dtype=np.float32)))
            uu, ii = np.meshgrid(np.arange(U), np.arange(I), indexing='ij')
            X = torch.from_numpy(np.stack([uu.reshape(-1), ii.reshape(-1)], 1).astype(np.int64))
            pred = model.predict({'X': X})['prediction'].detach().numpy().reshape(U, I)
            out = {'mf_user': model.uid_embeddings.weight.detach().numpy(), 'mf_item': model.iid_embeddings.weight.detach().numpy(),
                   'mf_user_bias': model.user_bias.weight.detach().numpy().reshape(-1),
                   'mf_item_bias': model.item_bias.weight.detach().numpy().reshape(-1),
                   'mf_global_bias': np.float32(model.global_bias.detach().numpy()), 'propensity': prop,
                   'mf_min_propensity': np.float32(M), 'pred': pred.astype(np.float32)}
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, pred.shape, float(np.abs(pred).max()))
    finally:
        shutil.rmtree(tmp)


def log_lines(messages, root):
    """The log of a run as comparable lines: the messages whose wording and numbers are part of the interface (data
    sizes, parameter count, optimizer, Init / Epoch / Best Iter lines with their %.4f metrics, early stop, model save /
    load), wall-clock durations masked, the temporary directory replaced by <root>."""
    import re
    keep = ('load ', 'size of ', 'label:', '# of ', 'Model # of', 'Drop Neg', 'Prepare ', 'Optimizer:', 'Init:', 'Epoch ',
            'Best Iter', 'Early stop', 'Save model', 'Load model', 'building ', 'loss = ', 'l2 inappropriate', 'Test Before', 'Test After', 'Save Test Results', '# cuda devices',
            'DataLoader:', 'Model:', 'Runner:', 'DataProcessor:')
    out = []
    for m in messages:
        m = m.strip()
        if m.startswith(keep):
            out.append(re.sub(r'\[\d+\.\d+ s\]', '[T s]', m.replace(root, '<root>')))
    return out


def make_run_fixture(name='run_recmodel', rank=1, seed=2019, epochs=3, n_users=120, n_items=150, per_user=10, test_neg_n=5,
                     batch_size=64, lr=0.01, l2=1e-4):
    """A WHOLE RUN of the unmodified reference, src/main.py's sequence (main.py:101-192) with its own DataLoader,
    DataProcessor, RecModel and BaseRunner on CPU: seeds -> load -> model + init_paras -> drop_neg -> "Test Before
    Training" -> runner.train (fit / evaluate / model selection / load best) -> "Test After Training" -> predictions.
    RecModel (plain matrix factorisation, src/models/RecModel.py) needs no CUDA, so every number is reproducible on CPU:
    the host mirror (loader, processor, runner, BaseModel/RecModel, rmse / mae) has to reproduce all of it."""
    ref = rh.load_reference()
    RefRecModel = sys.modules['models.RecModel'].RecModel
    tmp = tempfile.mkdtemp()
    try:
        synth.write_dataset(tmp, 'toy', n_users, n_items, per_user, feat_dim=64, seed=seed + 5)
        model_path = os.path.join(tmp, 'model', 'm.pt')
        os.makedirs(os.path.dirname(model_path))
        import logging
        messages = []

        class _Collect(logging.Handler):
            def emit(self, record):
                messages.append(record.getMessage())

        collector = _Collect(level=logging.INFO)
        logging.getLogger().addHandler(collector)
        old_level = logging.getLogger().level
        logging.getLogger().setLevel(logging.INFO)
        with rh.cpu_shims():
            torch.manual_seed(seed)
            np.random.seed(seed)
            dl = ref.DataLoader(path=tmp, dataset='toy', label='label', sep=',')
            model = RefRecModel(label_min=dl.label_min, label_max=dl.label_max, feature_num=0, user_num=dl.user_num,
                                item_num=dl.item_num, u_vector_size=64, i_vector_size=64, random_seed=seed,
                                model_path=model_path)
            model.apply(model.init_paras)
            if rank == 1:                           # main.py:157-158
                dl.drop_neg()
            dp = ref.DataProcessor(dl, model, rank=rank, test_neg_n=test_neg_n)
            runner = ref.BaseRunner(optimizer='Adam', learning_rate=lr, epoch=epochs, batch_size=batch_size,
                                    eval_batch_size=16384, dropout=0.2, l2=l2, metrics='rmse,mae', check_epoch=1,
                                    early_stop=1)
            before = runner.evaluate(model, dp.get_test_data(), dp)
            runner.train(model, dp, skip_eval=0)
            after = runner.evaluate(model, dp.get_test_data(), dp, write_rank=True)      # main.py:188-190
            rank_text = open(os.path.join(dl.path, ref.global_p.RANK_FILE_NAME)).read()
            pred = runner.predict(model, dp.get_test_data(), dp)
            sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        logging.getLogger().removeHandler(collector)
        logging.getLogger().setLevel(old_level)
        rank_lines = rank_text.strip().split(chr(10))
        rank_rows = np.array([[float(x) for x in ln.split(chr(9))] for ln in rank_lines[1:]])
        out = {'log': np.array(log_lines(messages, tmp)), 'rank_header': np.array(rank_lines[0]), 'rank_rows': rank_rows,
               'seed': seed, 'rank': rank, 'epochs': epochs, 'n_users': n_users, 'n_items': n_items, 'per_user': per_user,
               'test_neg_n': test_neg_n, 'batch_size': batch_size, 'lr': lr, 'l2': l2,
               'before': np.array(before, dtype=np.float64), 'after': np.array(after, dtype=np.float64),
               'train_results': np.array(runner.train_results, dtype=np.float64),
               'valid_results': np.array(runner.valid_results, dtype=np.float64),
               'test_results': np.array(runner.test_results, dtype=np.float64), 'pred': np.asarray(pred, np.float32)}
        for k, v in sd.items():
            out['sd_' + k] = v
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, 'before', before, 'valid', runner.valid_results, 'after', after)
    finally:
        shutil.rmtree(tmp)


def make_run_fixture_dccf(name='run_dccf', seed=2019, epochs=2, n_users=120, n_items=150, per_user=10, test_neg_n=20,
                          batch_size=64, lr=1e-3, l2=1e-4):
    """A whole run of the unmodified reference with DCCF itself, in the deterministic configuration --std 0 --dropout 0
    (no feature noise, no dropout: the only random inputs are the shuffles / negatives on the numpy generator and the
    confounder draws on the torch CPU generator, both reproducible).  The harness draws the (zero) noise from a
    private generator, as a GPU run of the reference would from the CUDA generator, so the CPU generator sees exactly
    the confounder draws (ref_harness.use_device_rng).  Sequence of src/main.py:101-192 with ndcg@5 / recall@5 /
    precision@5: before, per epoch (train rmse/mae, validation, test), after, predictions, checkpoint."""
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    try:
        d = synth.write_dataset(tmp, 'toy', n_users, n_items, per_user, feat_dim=64, seed=seed + 5)
        model_path = os.path.join(tmp, 'model', 'm.pt')
        os.makedirs(os.path.dirname(model_path))
        rh.use_device_rng(torch.Generator().manual_seed(12345))
        try:
            with rh.cpu_shims():
                torch.manual_seed(seed)
                np.random.seed(seed)
                dl = ref.DataLoader(path=tmp, dataset='toy', label='label', sep=',')
                model = rh.build_reference_model(ref, d, 'toy', SENT, dl.user_num, dl.item_num, std=0.0,
                                                 random_seed=seed, model_path=model_path)
                dl.drop_neg()
                dp = ref.DataProcessor(dl, model, rank=1, test_neg_n=test_neg_n)
                runner = ref.BaseRunner(optimizer='Adam', learning_rate=lr, epoch=epochs, batch_size=batch_size,
                                        eval_batch_size=16384, dropout=0.0, l2=l2,
                                        metrics='ndcg@5,recall@5,precision@5', check_epoch=1, early_stop=1)
                before = runner.evaluate(model, dp.get_test_data(), dp)
                runner.train(model, dp, skip_eval=0)
                after = runner.evaluate(model, dp.get_test_data(), dp)
                pred = runner.predict(model, dp.get_test_data(), dp)
                sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
                model.save_model()                   # the reference's own checkpoint writer (BaseModel.py:224-236)
                shutil.copy(model_path, os.path.join(GOLDEN, name + '_checkpoint.pt'))
        finally:
            rh.use_device_rng(None)
        out = {'seed': seed, 'epochs': epochs, 'n_users': n_users, 'n_items': n_items, 'per_user': per_user,
               'test_neg_n': test_neg_n, 'batch_size': batch_size, 'lr': lr, 'l2': l2,
               'before': np.array(before, dtype=np.float64), 'after': np.array(after, dtype=np.float64),
               'train_results': np.array(runner.train_results, dtype=np.float64),
               'valid_results': np.array(runner.valid_results, dtype=np.float64),
               'test_results': np.array(runner.test_results, dtype=np.float64), 'pred': np.asarray(pred, np.float32)}
        for k, v in sd.items():
            out['sd_' + k] = v
        np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
        print(name, 'before', before, 'valid', runner.valid_results, 'after', after)
    finally:
        shutil.rmtree(tmp)


def make_termination_fixture(name='termination', seed=11, n=300):
    """BaseRunner.eva_termination (src/runners/BaseRunner.py:193-210) and utils.best_result / format_metric
    (src/utils/utils.py:62-106) of the unmodified reference on random validation histories."""
    import json
    ref = rh.load_reference()
    rs = np.random.RandomState(seed)
    cases = []
    for _ in range(n):
        metric = ['ndcg@5', 'rmse'][rs.randint(2)]
        length = int(rs.randint(1, 45))
        kind = rs.randint(4)
        base = rs.random_sample(length)
        if kind == 1:
            base = np.sort(base)                                 # monotone up
        elif kind == 2:
            base = np.sort(base)[::-1]                           # monotone down
        elif kind == 3 and length > 25:
            base[int(rs.randint(0, length - 21))] = 2.0 if metric == 'ndcg@5' else -1.0   # an early best
        hist = [[float(np.round(v, 4)), float(np.round(rs.random_sample(), 4))] for v in base]
        runner = ref.BaseRunner(optimizer='Adam', learning_rate=1e-3, epoch=1, batch_size=8, eval_batch_size=8, dropout=0.0,
                                l2=0.0, metrics=metric + ',recall@5', check_epoch=1, early_stop=1)
        runner.valid_results = [list(h) for h in hist]
        cases.append({'metric': metric, 'history': hist, 'stop': bool(runner.eva_termination(None)),
                      'best': ref.utils.best_result(metric, [list(h) for h in hist]),
                      'fmt': ref.utils.format_metric(hist[-1])})
    with open(os.path.join(GOLDEN, name + '.json'), 'w') as f:
        json.dump(cases, f)
    print(name, n, 'cases,', sum(c['stop'] for c in cases), 'stops')


def make_cli_fixture(name='cli_flags'):
    """Every command-line flag the reference defines for the DCCF path — option strings, destination, default, type —
    from its own parsers in main.py's order (src/main.py:52-59): global, DataLoader, DCCF (-> DMF -> RecModel ->
    BaseModel), BaseRunner, DataProcessor."""
    import argparse
    import json
    ref = rh.load_reference()
    p = argparse.ArgumentParser()
    p = ref.utils.parse_global_args(p)
    p = ref.DataLoader.parse_data_args(p)
    p = ref.DCCF.parse_model_args(p, model_name='DCCF')
    p = ref.BaseRunner.parse_runner_args(p)
    p = ref.DataProcessor.parse_dp_args(p)
    flags = [{'options': list(a.option_strings), 'dest': a.dest, 'default': a.default,
              'type': getattr(a.type, '__name__', None)} for a in p._actions if a.dest != 'help']
    with open(os.path.join(GOLDEN, name + '.json'), 'w') as f:
        json.dump(flags, f, indent=1)
    print(name, len(flags), 'flags')


MAIN_ARGS = ['--rank', '1', '--model_name', 'RecModel', '--optimizer', 'Adam', '--lr', '0.01', '--dataset', 'toy',
             '--metric', 'rmse,mae', '--gpu', '', '--epoch', '2', '--batch_size', '64', '--test_neg_n', '5',
             '--random_seed', '7']


def make_main_fixture(name='main_recmodel'):
    """The reference's src/main.py ITSELF, run as a script (RecModel on CPU, default log / model / result locations
    relative to the working directory `<tmp>/src`): which files it creates — their names are derived from the
    hyper-parameters (main.py:63-84) —, what it logs and what it saves as the result."""
    import json
    import logging
    import runpy
    ref = rh.load_reference()
    tmp = tempfile.mkdtemp()
    cwd = os.getcwd()
    argv = list(sys.argv)
    try:
        synth.write_dataset(os.path.join(tmp, 'datasets'), 'toy', 120, 150, 10, feat_dim=64, seed=9)
        os.makedirs(os.path.join(tmp, 'src'))
        os.makedirs(os.path.join(tmp, 'result'))          # the reference never creates it (breakage 5)
        os.chdir(os.path.join(tmp, 'src'))
        sys.argv = ['main.py'] + MAIN_ARGS + ['--path', '../datasets/']
        sys.path.insert(0, rh.REFERENCE_SRC)
        try:
            with rh.cpu_shims():
                runpy.run_path(os.path.join(rh.REFERENCE_SRC, 'main.py'), run_name='__main__')
        finally:
            sys.path.remove(rh.REFERENCE_SRC)
            for h in logging.root.handlers[:]:
                logging.root.removeHandler(h)
                h.close()
        files = sorted(os.path.relpath(os.path.join(r, f), tmp) for sub in ('log', 'model', 'result')
                       for r, _, fs in os.walk(os.path.join(tmp, sub)) for f in fs)
        log_path = [f for f in files if f.startswith('log')][0]
        res_path = [f for f in files if f.startswith('result')][0]
        import re
        messages = [re.sub(r'^(INFO|WARNING|ERROR|DEBUG):root:', '', ln)
                    for ln in open(os.path.join(tmp, log_path)).read().split(chr(10))]
        out = {'args': MAIN_ARGS, 'files': files, 'log': log_lines(messages, tmp),
               'result': [float(x) for x in np.load(os.path.join(tmp, res_path))]}
        with open(os.path.join(GOLDEN, name + '.json'), 'w') as f:
            json.dump(out, f)
        print(name, files)
    finally:
        os.chdir(cwd)
        sys.argv = argv
        shutil.rmtree(tmp)


def make_metrics_fixture(name='metrics', seed=5):
    ref = rh.load_reference()
    rs = np.random.RandomState(seed)
    uid, iid, Y = [], [], []
    for u in range(25):
        n_pos = rs.randint(1, 4)
        n = n_pos + 50
        items = rs.choice(500, n, replace=False)
        uid += [u * 3] * n
        iid += list(items)
        Y += [1.0] * n_pos + [0.0] * 50
    uid, iid, Y = np.array(uid), np.array(iid), np.array(Y, dtype=np.float32)
    perm = rs.permutation(len(uid))
    uid, iid, Y = uid[perm], iid[perm], Y[perm]
    p = (rs.standard_normal(len(uid)) + 0.8 * Y).astype(np.float32)      # continuous -> no ties
    data = {'uid': uid, 'iid': iid, 'Y': Y}
    metrics = ['ndcg@5', 'hit@5', 'precision@5', 'recall@5', 'f1@5', 'ndcg@10', 'recall@1', 'precision@20']
    vals = ref.BaseModel.evaluate_method(p, data, metrics)
    kat = {
        'ndcg_2120_k4_m1': ref.rank_metrics.ndcg_at_k([2, 1, 2, 0], 4, method=1),
        'dcg_k2_m1': ref.rank_metrics.dcg_at_k([3, 2, 3, 0, 0, 1, 2, 2, 3, 0], 2, method=1),
        'prec_001_k3': ref.rank_metrics.precision_at_k([0, 0, 1], 3),
        'ndcg_0_k1': ref.rank_metrics.ndcg_at_k([0], 1, method=1),
        'ndcg_1_k2': ref.rank_metrics.ndcg_at_k([1], 2, method=1),
    }
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), uid=uid, iid=iid, Y=Y, p=p, metrics=np.array(metrics),
                        values=np.array([float(v) for v in vals], dtype=np.float64),
                        **{'kat_' + k: np.float64(v) for k, v in kat.items()})
    print(name, dict(zip(metrics, vals)), kat)


if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    if '--skip-train' not in sys.argv:
      make_train_fixture('train_f64', U=40, I=50, F=64, P=16, steps=3)
      make_train_fixture('train_f768', U=24, I=30, F=768, P=8, steps=2)
      make_train_fixture('train_nodrop', U=40, I=50, F=64, P=12, steps=2, dropout=0.0, std=0.0, S=4, A=3)
    make_sampler_fixture()
    make_config0_digest()
    make_config0_train()
    make_config0_eval()
    make_ipsmf_fixture()
    make_run_fixture()
    make_run_fixture(name='run_recmodel_rank0', rank=0)
    make_run_fixture_dccf()
    make_termination_fixture()
    make_cli_fixture()
    make_main_fixture()
    make_metrics_fixture()
