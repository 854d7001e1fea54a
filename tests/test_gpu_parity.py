"""Parity of the CUDA path (through the C-ABI) with the reference fixtures and the oracle.  Needs a GPU."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_params, rel_err
from oracle import dccf_oracle as O

pytestmark = pytest.mark.gpu

TRAIN_FIXTURES = ['train_f64', 'train_f768', 'train_nodrop']
# updated embeddings: 1e-5 relative (north_star) on the default-configuration fixtures.  train_nodrop (std 0,
# dropout 0, A = 3 identical attribute copies) has touched item rows whose gradient nearly cancels the l2 term,
# |g| ~ eps = 1e-8, where Adam amplifies fp32 summation noise by lr/eps; the numpy oracle itself is 6e-6 away
# from the reference there.  Its exp_avg / exp_avg_sq (linear / quadratic in g) are still held to 2e-5.
EMB_TOL = {'train_f64': 1e-5, 'train_f768': 1e-5, 'train_nodrop': 5e-5}


def make_model(params, S, A, std, seed=2019, expo_factors=None):
    from dccf_b200.models.DCCF import DCCF
    U, I = params['E_user'].shape[0], params['E_item'].shape[0]
    model = DCCF(path='', dataset='', sentence_model='', sample_num=S, attribute_num=A, std=std, label_min=0,
                 label_max=1, feature_num=0, user_num=U, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1,
                 random_seed=seed, model_path='/tmp/dccf_test_model.pt', feature_embedding=params['Feat'],
                 expo_prob=None if expo_factors is not None else params['expo'], expo_factors=expo_factors)
    with torch.no_grad():
        model.uid_embeddings.weight.copy_(torch.from_numpy(params['E_user']))
        model.iid_embeddings.weight.copy_(torch.from_numpy(params['E_item']))
        model.mlp[0].weight.copy_(torch.from_numpy(params['W']))
        model.mlp[0].bias.copy_(torch.from_numpy(params['b']))
    return model.cuda()


def feed(g, t, drop, rank=1, train=True):
    X = g['X_%d' % t]
    b = X.shape[0] // 2
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': rank, 'train': train, 'dropout': drop,
          'Y': torch.cat([torch.ones(b), torch.zeros(X.shape[0] - b)]).cuda(),
          'sample_item': torch.from_numpy(g['sample_item_%d' % t]), 'noise': torch.from_numpy(g['noise_%d' % t])}
    if drop > 0:
        fd['dropout_mask'] = torch.from_numpy(g['mask_%d' % t])
    return fd


def model_params(model):
    return {'E_user': model.uid_embeddings.weight.detach().cpu().numpy(),
            'E_item': model.iid_embeddings.weight.detach().cpu().numpy(),
            'W': model.mlp[0].weight.detach().cpu().numpy(), 'b': model.mlp[0].bias.detach().cpu().numpy()}


# ---------------------------------------------------------------------------------------------------------
# against the reference's own outputs (golden fixtures)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_predict_matches_reference(golden, name):
    g = golden(name)
    drop = float(g['dropout'])
    model = make_model(golden_params(g), int(g['S']), int(g['A']), float(g['std']))
    out = model.predict(feed(g, 0, drop, train=False))
    assert rel_err(out['prediction'].cpu().numpy(), g['pred_0']) < 1e-5          # 1e-5 relative, fp32
    model.check_ids()


@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_eval_predict_matches_reference(golden, name):
    """Ragged batch (37 pairs), dropout off, after training (final weights)."""
    g = golden(name)
    model = make_model(golden_params(g, 'final_'), int(g['S']), int(g['A']), float(g['std']))
    fd = {'X': torch.from_numpy(g['eval_X']).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(g['eval_sample_item']), 'noise': torch.from_numpy(g['eval_noise'])}
    out = model.predict(fd)
    assert rel_err(out['prediction'].cpu().numpy(), g['eval_pred']) < 1e-5


def _check_trained(model, opt_state, g, emb_tol=1e-5):
    got = model_params(model)
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert rel_err(opt_state['m'][k], g['m_' + k]) < 2e-5, k
        assert rel_err(opt_state['v'][k], g['v_' + k]) < 2e-5, k
    # see tests/test_oracle_golden.py for why W gets a looser bound (lr/eps conditioning of Adam where |g| << eps)
    assert rel_err(got['E_user'], g['final_E_user']) < emb_tol
    assert rel_err(got['E_item'], g['final_E_item']) < emb_tol
    assert rel_err(got['b'], g['final_b']) < emb_tol
    assert rel_err(got['W'], g['final_W']) < 5e-4


@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_fused_training_matches_reference(golden, name):
    """model.train_step x steps == reference BaseRunner.fit (forward, BPR, l2, backward, clip, Adam)."""
    g = golden(name)
    drop, steps = float(g['dropout']), int(g['steps'])
    model = make_model(golden_params(g), int(g['S']), int(g['A']), float(g['std']))
    model.optimizer = model.make_fused_optimizer(lr=float(g['lr']), l2=float(g['l2']))
    for t in range(steps):
        out = model.train_step(feed(g, t, drop))
        assert rel_err(out['prediction'].cpu().numpy(), g['pred_%d' % t]) < 2e-5
        assert abs(float(out['loss']) - float(g['loss_%d' % t])) < 1e-5 * abs(float(g['loss_%d' % t]))
    opt = model.optimizer
    _check_trained(model, {'m': {k: v.cpu().numpy() for k, v in opt.exp_avg.items()},
                           'v': {k: v.cpu().numpy() for k, v in opt.exp_avg_sq.items()}}, g, EMB_TOL[name])
    assert int((opt.head_u != -1).sum()) == 0 and int((opt.head_i != -1).sum()) == 0
    model.check_ids()


@pytest.mark.parametrize('name', ['train_f64', 'train_nodrop'])
def test_autograd_training_matches_reference(golden, name):
    """The reference's own sequence — loss + l2*model.l2(), backward, clip_grad_value_, torch Adam — on top of
    the CUDA forward/backward kernels (drop-in for an unmodified runner)."""
    g = golden(name)
    drop, steps = float(g['dropout']), int(g['steps'])
    model = make_model(golden_params(g), int(g['S']), int(g['A']), float(g['std']))
    optim = torch.optim.Adam(model.parameters(), lr=float(g['lr']), weight_decay=float(g['l2']))
    model.train()
    for t in range(steps):
        optim.zero_grad()
        out = model(feed(g, t, drop))
        loss = out['loss'] + model.l2() * float(g['l2'])
        loss.backward()
        torch.nn.utils.clip_grad_value_(model.parameters(), 50)
        optim.step()
        assert abs(float(out['loss'].detach()) - float(g['loss_%d' % t])) < 1e-5 * abs(float(g['loss_%d' % t]))
    plist = {'E_user': model.uid_embeddings.weight, 'E_item': model.iid_embeddings.weight,
             'W': model.mlp[0].weight, 'b': model.mlp[0].bias}
    for k, p in plist.items():
        assert rel_err(p.grad.cpu().numpy(), g['lastgrad_' + k]) < 2e-5, k
    _check_trained(model, {'m': {k: optim.state[p]['exp_avg'].cpu().numpy() for k, p in plist.items()},
                           'v': {k: optim.state[p]['exp_avg_sq'].cpu().numpy() for k, p in plist.items()}}, g,
                   EMB_TOL[name])


def test_ranker_matches_reference_metrics(golden):
    from dccf_b200.models.BaseModel import BaseModel
    g = golden('metrics')
    data = {'uid': g['uid'], 'iid': g['iid'], 'Y': g['Y']}
    vals = BaseModel.evaluate_method(g['p'], data, [str(m) for m in g['metrics']])
    assert np.abs(np.array(vals) - g['values']).max() < 1e-6                      # metrics to 1e-6


# ---------------------------------------------------------------------------------------------------------
# against the oracle on seeded inputs
# ---------------------------------------------------------------------------------------------------------
def random_problem(seed, U, I, F, P, S, A, std, drop):
    rs = np.random.RandomState(seed)
    params = {'E_user': (rs.standard_normal((U, 64)) * 0.05).astype(np.float32),
              'E_item': (rs.standard_normal((I, 64)) * 0.05).astype(np.float32),
              'W': (rs.standard_normal((64, 64 + F)) * 0.05).astype(np.float32),
              'b': (rs.standard_normal(64) * 0.05).astype(np.float32),
              'Feat': (rs.standard_normal((I, F)) / np.sqrt(F)).astype(np.float32),
              'expo': rs.random_sample((U, I)).astype(np.float32)}
    b = P // 2
    u = rs.randint(0, U, size=b)
    X = np.concatenate([np.stack([u, rs.randint(0, I, size=b)], 1), np.stack([u, rs.randint(0, I, size=b)], 1)])
    if P % 2:
        X = np.concatenate([X, [[rs.randint(0, U), rs.randint(0, I)]]])
    X = X.astype(np.int64)
    si = rs.randint(0, I, size=(P, S)).astype(np.int64)
    N = P * (S + 1) * A
    noise = (rs.standard_normal((N, F)) * std).astype(np.float32) if std > 0 else None
    mask = ((rs.random_sample((N, 64)) < 1 - drop) / (1 - drop)).astype(np.float32) if drop > 0 else None
    return params, X, si, noise, mask


@pytest.mark.parametrize('U,I,F,P,S,A,std,drop', [
    (300, 500, 768, 256, 10, 2, 0.1, 0.2),     # the reference's training step shape
    (50, 70, 128, 37, 3, 1, 0.1, 0.0),         # ragged: 37*4 rows is not a multiple of the 128-row tile
    (20, 30, 64, 2, 0, 1, 0.0, 0.0),           # no confounders: single slot
    (64, 64, 256, 130, 5, 3, 0.0, 0.5),
])
def test_forward_backward_vs_oracle(U, I, F, P, S, A, std, drop):
    params, X, si, noise, mask = random_problem(11, U, I, F, P, S, A, std, drop)
    Xe = X[:P - (P % 2)]
    sie = si[:len(Xe)]
    R = (S + 1) * A
    noise_e = noise[:len(Xe) * R] if noise is not None else None
    mask_e = mask[:len(Xe) * R] if mask is not None else None
    model = make_model(params, S, A, std)
    # forward on the full (possibly odd) batch
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': drop,
          'sample_item': torch.from_numpy(si)}
    if noise is not None:
        fd['noise'] = torch.from_numpy(noise)
    if mask is not None:
        fd['dropout_mask'] = torch.from_numpy(mask)
    ref = O.predict(params, X, si, noise, mask, A, dtype=np.float64)
    got = model.predict(fd)['prediction'].cpu().numpy()
    assert rel_err(got, ref['pred']) < 1e-5
    # gradients through the autograd path on the even part
    fd = {'X': torch.from_numpy(Xe).cuda(), 'rank': 1, 'train': True, 'dropout': drop,
          'Y': torch.zeros(len(Xe)).cuda(), 'sample_item': torch.from_numpy(sie)}
    if noise_e is not None:
        fd['noise'] = torch.from_numpy(noise_e)
    if mask_e is not None:
        fd['dropout_mask'] = torch.from_numpy(mask_e)
    model.train()
    out = model(fd)
    out['loss'].backward()
    fwd = O.predict(params, Xe, sie, noise_e, mask_e, A, dtype=np.float64)
    grads = O.backward(params, Xe, sie, noise_e, mask_e, A, fwd, dtype=np.float64)
    assert abs(float(out['loss']) - O.loss_bpr(fwd['pred'])) < 1e-5 * abs(O.loss_bpr(fwd['pred']))
    assert rel_err(model.uid_embeddings.weight.grad.cpu().numpy(), grads['E_user']) < 1e-5
    assert rel_err(model.iid_embeddings.weight.grad.cpu().numpy(), grads['E_item']) < 1e-5
    assert rel_err(model.mlp[0].weight.grad.cpu().numpy(), grads['W']) < 1e-5
    assert rel_err(model.mlp[0].bias.grad.cpu().numpy(), grads['b']) < 1e-5
    model.check_ids()


def test_mse_loss_mode_vs_oracle():
    """rank == 0: MSELoss(prediction, Y) (src/models/DCCF.py:123-125) through the fused step."""
    U, I, F, P, S, A = 40, 60, 64, 24, 4, 2
    params, X, si, noise, mask = random_problem(5, U, I, F, P, S, A, 0.1, 0.2)
    Y = np.random.RandomState(1).random_sample(P).astype(np.float32)
    model = make_model(params, S, A, 0.1)
    model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 0, 'train': True, 'dropout': 0.2, 'Y': torch.from_numpy(Y).cuda(),
          'sample_item': torch.from_numpy(si), 'noise': torch.from_numpy(noise), 'dropout_mask': torch.from_numpy(mask)}
    out = model.train_step(fd)
    state = {k: {'m': np.zeros_like(params[k]), 'v': np.zeros_like(params[k])} for k in ('E_user', 'E_item', 'W', 'b')}
    hp = dict(lr=1e-3, l2=1e-4, weight_decay=1e-4)
    p2, s2, loss, pred = O.train_step(params, state, 1, X, si, noise, mask, A, hp, loss_mode=1, Y=Y)
    assert abs(float(out['loss']) - loss) < 1e-5 * abs(loss)
    opt = model.optimizer
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert rel_err(opt.exp_avg[k].cpu().numpy(), s2[k]['m']) < 2e-5, k


def test_ipsmf_exposure_vs_oracle():
    """Exposure computed on the fly from IPSBiasedMF factors (src/models/IPSBiasedMF.py:42-53) instead of the
    dense user x item matrix (scaled config)."""
    from dccf_b200 import synth
    U, I, F, P, S, A = 80, 120, 64, 64, 10, 2
    params, X, si, noise, mask = random_problem(7, U, I, F, P, S, A, 0.1, 0.0)
    fac = synth.make_ipsmf_factors(U, I, seed=3)
    model = make_model(params, S, A, 0.1, expo_factors=fac)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(si), 'noise': torch.from_numpy(noise)}
    got = model.predict(fd)['prediction'].cpu().numpy()
    ref = O.predict(params, X, si, noise, None, A, dtype=np.float64, expo=fac)
    assert rel_err(got, ref['pred']) < 1e-5


@pytest.mark.parametrize('dense_expo', [False, True])
def test_row_sharded_user_table_equals_full_table(dense_expo):
    """Scaled configuration: a rank that owns users [lo, hi) holds only those rows of the user
    table, of its Adam state and of the exposure source (IPS-MF user factors, or the dense rows), and is fed
    batches of its own users with GLOBAL ids.  Three fused steps give the same predictions, loss, user rows,
    item table and W as the unsharded model on the same batches."""
    from dccf_b200 import synth
    from dccf_b200.models.DCCF import DCCF
    U, I, F, P, S, A = 400, 300, 128, 64, 10, 2
    lo, hi = 150, 270
    params, X, si, _, _ = random_problem(29, U, I, F, P, S, A, 0.0, 0.0)
    rs = np.random.RandomState(4)
    X[:P // 2, 0] = rs.randint(lo, hi, size=P // 2)
    X[P // 2:, 0] = X[:P // 2, 0]
    fac = None if dense_expo else synth.make_ipsmf_factors(U, I, seed=3)
    outs = []
    for shard in (None, (lo, hi)):
        model = DCCF(path='', dataset='', sentence_model='', sample_num=S, attribute_num=A, std=0.1, label_min=0,
                     label_max=1, feature_num=0, user_num=U, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1,
                     random_seed=2019, model_path='/tmp/dccf_test_model.pt', feature_embedding=params['Feat'],
                     expo_prob=params['expo'] if dense_expo else None, expo_factors=fac, user_shard=shard)
        with torch.no_grad():
            eu = params['E_user'] if shard is None else params['E_user'][lo:hi]
            model.uid_embeddings.weight.copy_(torch.from_numpy(eu))
            model.iid_embeddings.weight.copy_(torch.from_numpy(params['E_item']))
            model.mlp[0].weight.copy_(torch.from_numpy(params['W']))
            model.mlp[0].bias.copy_(torch.from_numpy(params['b']))
        model = model.cuda()
        assert model.uid_embeddings.weight.shape[0] == (U if shard is None else hi - lo)
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        preds, losses = [], []
        for t in range(3):
            Xt = np.roll(X, t, axis=0).copy()
            Xt[P // 2:, 0] = Xt[:P // 2, 0]
            o = model.train_step({'X': torch.from_numpy(Xt).cuda(), 'rank': 1, 'train': True, 'dropout': 0.2,
                                  'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(np.roll(si, t, axis=0).copy())})
            preds.append(o['prediction'].cpu().numpy().copy())
            losses.append(float(o['loss']))
        model.check_ids()
        outs.append((preds, losses, model_params(model)))
    full, part = outs
    for t in range(3):
        assert np.array_equal(part[0][t], full[0][t])
        assert part[1][t] == full[1][t]
    assert np.array_equal(part[2]['E_user'], full[2]['E_user'][lo:hi])
    for k in ('E_item', 'W', 'b'):
        assert np.array_equal(part[2][k], full[2][k]), k
    # a user outside the shard is reported, not silently mapped onto a local row
    bad = X.copy()
    bad[0, 0] = bad[P // 2, 0] = lo - 1
    model.predict({'X': torch.from_numpy(bad).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
                   'sample_item': torch.from_numpy(si)})
    with pytest.raises(IndexError):
        model.check_ids()


def test_training_state_resume_is_bit_identical(tmp_path):
    """save_training_state / load_training_state (weights, Adam moments, step, stream position): a run resumed
    in a fresh model continues exactly like the uninterrupted one."""
    U, I, F, P, S, A = 120, 150, 128, 32, 10, 2
    params, X, si, _, _ = random_problem(31, U, I, F, P, S, A, 0.0, 0.0)
    X[P // 2:, 0] = X[:P // 2, 0]

    def steps(model, t0, t1):
        out = []
        for t in range(t0, t1):
            o = model.train_step({'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': True, 'dropout': 0.2,
                                  'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(np.roll(si, t, axis=0).copy())})
            out.append(float(o['loss']))
        return out

    a = make_model(params, S, A, 0.1)
    a.optimizer = a.make_fused_optimizer(lr=1e-3, l2=1e-4)
    steps(a, 0, 3)
    path = a.save_training_state(str(tmp_path / 'run.train_state'))
    la = steps(a, 3, 6)
    b = make_model(params, S, A, 0.1)
    b.load_training_state(path)
    lb = steps(b, 3, 6)
    assert la == lb
    pa, pb = model_params(a), model_params(b)
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
    for k in a.optimizer.exp_avg:
        assert torch.equal(a.optimizer.exp_avg[k], b.optimizer.exp_avg[k])
        assert torch.equal(a.optimizer.exp_avg_sq[k], b.optimizer.exp_avg_sq[k])


# ---------------------------------------------------------------------------------------------------------
# the library's own random streams
# ---------------------------------------------------------------------------------------------------------
def test_fused_rng_equals_materialised_rng():
    """rng mode 2 (Philox in registers) is bit-identical to mode 1 fed with dccf_noise_fill /
    dccf_dropout_mask_fill output for the same (seed, offset): forward AND gradients."""
    from dccf_b200 import kernels
    U, I, F, P, S, A, std, drop = 100, 150, 768, 64, 10, 2, 0.1, 0.2
    params, X, si, _, _ = random_problem(3, U, I, F, P, S, A, 0.0, 0.0)
    N = P * (S + 1) * A
    seed = 2019
    results = []
    for explicit in (False, True):
        model = make_model(params, S, A, std, seed=seed)
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': True, 'dropout': drop,
              'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(si)}
        if explicit:
            noise = torch.empty((N, F), device='cuda')
            mask = torch.empty((N, 64), device='cuda')
            kernels.noise_fill(noise, std, seed, 1)          # first predict call of a fresh model -> offset 1
            kernels.dropout_mask_fill(mask, drop, seed, 1)
            fd['noise'], fd['dropout_mask'] = noise, mask
        out = model.train_step(fd)
        results.append((out['prediction'].cpu().numpy(), model_params(model)))
    assert np.array_equal(results[0][0], results[1][0])
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert np.array_equal(results[0][1][k], results[1][1][k]), k


def test_noise_and_mask_statistics():
    from dccf_b200 import kernels
    noise = torch.empty((4096, 768), device='cuda')
    kernels.noise_fill(noise, 0.1, 2019, 5)
    n = noise.double()
    assert abs(float(n.mean())) < 2e-4 and abs(float(n.std()) - 0.1) < 2e-4
    z = n / 0.1
    assert abs(float((z ** 3).mean())) < 0.01 and abs(float((z ** 4).mean()) - 3.0) < 0.02
    assert abs(float((n[:, ::2] * n[:, 1::2]).mean())) < 1e-5            # neighbouring columns uncorrelated
    other = torch.empty_like(noise)
    kernels.noise_fill(other, 0.1, 2019, 6)
    assert abs(float((n * other.double()).mean())) < 1e-5                # calls are independent
    mask = torch.empty((4096, 64), device='cuda')
    kernels.dropout_mask_fill(mask, 0.2, 2019, 5)
    vals = torch.unique(mask).cpu().numpy()
    assert np.allclose(vals, [0.0, 1.25])
    assert abs(float((mask > 0).double().mean()) - 0.8) < 5e-3
    # row0 offsets address the same stream
    part = torch.empty((100, 768), device='cuda')
    kernels.noise_fill(part, 0.1, 2019, 5, row0=1000)
    assert torch.equal(part, noise[1000:1100])


# ---------------------------------------------------------------------------------------------------------
# ranker: ties, ragged and tiny users, top-k ids bit-exact
# ---------------------------------------------------------------------------------------------------------
def test_ranker_topk_ids_bit_exact_with_ties():
    from dccf_b200.models.BaseModel import group_candidates, rank_metrics_device
    rs = np.random.RandomState(4)
    uid, iid, Y = [], [], []
    for u in range(200):
        n = int(rs.choice([1, 2, 4, 5, 37, 1001, 1003]))
        uid += [u] * n
        iid += list(rs.choice(5000, n, replace=False))
        Y += list((rs.random_sample(n) < 0.1).astype(np.float32))
        Y[-1] = 1.0
    uid, iid, Y = np.array(uid), np.array(iid, dtype=np.int64), np.array(Y, dtype=np.float32)
    perm = rs.permutation(len(uid))
    uid, iid, Y = uid[perm], iid[perm], Y[perm]
    scores = np.round(rs.standard_normal(len(uid)), 1).astype(np.float32)       # coarse grid -> many ties
    scores[rs.randint(0, len(uid), 20)] = np.nan
    for k in (1, 5, 10):
        users, want_ids, want_rows, want_m = O.rank_users(scores, uid, Y, iid, k)
        _, rows, off = group_candidates(uid)
        m, topk = rank_metrics_device(torch.from_numpy(scores).cuda(), torch.from_numpy(Y).cuda(),
                                      torch.from_numpy(iid).cuda(), torch.from_numpy(rows).cuda(),
                                      torch.from_numpy(off).cuda(), k, want_topk=True)
        assert np.array_equal(topk.cpu().numpy(), want_ids)                      # bit-exact ids, ties by item id
        assert np.abs(m.cpu().numpy() - want_m).max() < 1e-12


def test_ranker_all_metrics_one_launch_grouped_rows_and_graded_labels():
    """dccf_rank_eval_multi: every metric at several k from ONE launch (per-user values and their fixed-order sums),
    the contiguous path (rows already grouped by user: no cand_rows indirection), graded (non 0/1) labels through the
    ideal-DCG fallback, and k > 16 through the selection kernel — all against the oracle's per-user loop."""
    from dccf_b200 import kernels
    from dccf_b200.models.BaseModel import BaseModel, group_candidates, rank_metrics_device, rank_sums_device
    rs = np.random.RandomState(11)
    for graded in (False, True):
        uid, iid, Y = [], [], []
        for u in range(300):
            n = int(rs.choice([1, 3, 8, 16, 17, 33, 257, 1001, 1004]))
            uid += [u] * n
            iid += list(rs.choice(6000, n, replace=False))
            lab = (rs.random_sample(n) < 0.15).astype(np.float32)
            if graded:
                lab = lab * rs.randint(1, 6, size=n).astype(np.float32)
            lab[rs.randint(n)] = 1.0 if not graded else 3.0
            Y += list(lab)
        uid, iid, Y = np.array(uid), np.array(iid, dtype=np.int64), np.array(Y, dtype=np.float32)
        scores = np.round(rs.standard_normal(len(uid)), 2).astype(np.float32)
        scores[rs.randint(0, len(uid), 30)] = np.nan
        for shuffled in (False, True):
            if shuffled:
                perm = rs.permutation(len(uid))
                uid, iid, Y, scores = uid[perm], iid[perm], Y[perm], scores[perm]
            _, rows, off = group_candidates(uid)
            grouped = np.array_equal(rows, np.arange(len(rows)))
            assert grouped == (not shuffled)
            rows_d = None if grouped else torch.from_numpy(rows).cuda()
            sc, lb, ii, of = (torch.from_numpy(scores).cuda(), torch.from_numpy(Y).cuda(), torch.from_numpy(iid).cuda(),
                              torch.from_numpy(off).cuda())
            ks = [1, 5, 10, 16]
            n_users = len(off) - 1
            per_user = torch.empty((n_users, len(ks), 5), dtype=torch.float64, device='cuda')
            sums = torch.empty((len(ks), 5), dtype=torch.float64, device='cuda')
            topk = torch.empty((n_users, ks[-1]), dtype=torch.int64, device='cuda')
            before = kernels.LAUNCHES[0]
            kernels.rank_eval_multi(sc, lb, ii, rows_d, of, ks, out_metrics=per_user, out_sums=sums, out_topk_iid=topk)
            assert kernels.LAUNCHES[0] - before == 1
            per_user, sums = per_user.cpu().numpy(), sums.cpu().numpy()
            for j, k in enumerate(ks):
                _, want_ids, _, want_m = O.rank_users(scores, uid, Y, iid, k)
                assert np.array_equal(topk.cpu().numpy()[:, :k], want_ids)
                assert np.nanmax(np.abs(per_user[:, j] - want_m)) < 1e-12
                assert np.array_equal(np.isnan(per_user[:, j]), np.isnan(want_m))
                assert np.abs(sums[j] - want_m.sum(axis=0)).max() < 1e-9
            # a second call reuses the workspace (the kernel leaves its counter at zero)
            again = rank_sums_device(sc, lb, ii, rows_d, of, [10, 5]).cpu().numpy()
            assert np.array_equal(again[0], sums[2]) and np.array_equal(again[1], sums[1])
            # k > 16: the selection kernel
            _, want_ids, _, want_m = O.rank_users(scores, uid, Y, iid, 20)
            m20, t20 = rank_metrics_device(sc, lb, ii, rows_d, of, 20, want_topk=True)
            assert np.array_equal(t20.cpu().numpy(), want_ids)
            assert np.nanmax(np.abs(m20.cpu().numpy() - want_m)) < 1e-12
        data = {'uid': uid, 'iid': iid, 'Y': Y}
        names = ['ndcg@5', 'recall@5', 'precision@5', 'hit@10', 'f1@3', 'ndcg@20']
        before = kernels.LAUNCHES[0]
        got = BaseModel.evaluate_method(np.nan_to_num(scores, nan=-1e9), data, names)
        assert kernels.LAUNCHES[0] - before == 2            # one launch for k in {3, 5, 10}, one for k = 20
        want = O.evaluate_method(np.nan_to_num(scores, nan=-1e9), data, names)
        assert np.abs(np.array(got) - np.array(want)).max() < 1e-12


# ---------------------------------------------------------------------------------------------------------
# optimizer sweep
# ---------------------------------------------------------------------------------------------------------
def test_adam_sweep_duplicates_and_determinism():
    """Many records on one row, records on no row, result independent of atomics timing."""
    from dccf_b200 import kernels
    rs = np.random.RandomState(9)
    rows, n_rec = 1000, 700
    keys = rs.randint(0, rows, size=n_rec).astype(np.int32)
    keys[:300] = 17                                                              # 300 records on the same row
    grads = rs.standard_normal((n_rec, 64)).astype(np.float32) * 0.01
    table = rs.standard_normal((rows, 64)).astype(np.float32) * 0.05
    dense = np.zeros((rows, 64), dtype=np.float32)
    for r in range(n_rec):                                                       # ascending record order, fp32
        dense[keys[r]] += grads[r]
    want_p, want_m, want_v = O.adam_step(table, dense, np.zeros_like(table), np.zeros_like(table), 1)
    outs = []
    for _ in range(2):
        t = torch.from_numpy(table).cuda()
        m = torch.zeros_like(t)
        v = torch.zeros_like(t)
        head = torch.full((rows,), -1, dtype=torch.int32, device='cuda')
        nxt = torch.empty(n_rec, dtype=torch.int32, device='cuda')
        hp = kernels.make_adam(1e-3, 1e-4, 1e-4, step=1)
        kernels.adam_sweep(t, m, v, torch.from_numpy(keys).cuda(), torch.from_numpy(grads).cuda(), n_rec, head, nxt, hp)
        assert int((head != -1).sum()) == 0
        outs.append((t.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy()))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    assert rel_err(outs[0][1], want_m) < 1e-6 and rel_err(outs[0][2], want_v) < 1e-6
    assert rel_err(outs[0][0], want_p) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# full-size properties and edge cases
# ---------------------------------------------------------------------------------------------------------
def test_eval_batch_full_size_properties():
    """eval_batch_size = 16384 pairs (360 448 predictor rows): deterministic, finite, linear in the user
    embedding (s = <E_user, h>) and invariant to a constant shift of the exposure row (softmax)."""
    U, I, F, P, S, A = 2000, 5000, 768, 16384, 10, 2
    params, X, si, _, _ = random_problem(21, U, I, F, P, S, A, 0.0, 0.0)
    model = make_model(params, S, A, 0.1)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(si)}
    model._rng_offset = 0
    a = model.predict(fd)['prediction']
    model._rng_offset = 0
    b = model.predict(fd)['prediction']
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    c = model.predict(fd)['prediction']                       # next call -> new noise
    assert not torch.equal(a, c)
    with torch.no_grad():
        model.uid_embeddings.weight.mul_(2.0)
        model.expo_prob.add_(3.0)
    model._rng_offset = 0
    d = model.predict(fd)['prediction']
    assert rel_err(d.cpu().numpy(), 2.0 * a.cpu().numpy()) < 1e-5
    # spot check against the oracle with the materialised noise of the first 64 pairs
    from dccf_b200 import kernels
    R = (S + 1) * A
    noise = torch.empty((64 * R, F), device='cuda')
    kernels.noise_fill(noise, 0.1, model.random_seed, 1)
    p2 = dict(params)
    p2['E_user'] = params['E_user'] * 2
    ref = O.predict(p2, X[:64], si[:64], noise.cpu().numpy(), None, A, dtype=np.float64)
    assert rel_err(d[:64].cpu().numpy(), ref['pred']) < 1e-5


def test_out_of_range_ids_are_reported():
    params, X, si, _, _ = random_problem(2, 30, 40, 64, 8, 2, 1, 0.0, 0.0)
    model = make_model(params, 2, 1, 0.0)
    X = X.copy()
    X[3, 1] = 40
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(si)}
    model.predict(fd)
    with pytest.raises(IndexError):
        model.check_ids()


def test_empty_batch():
    params, X, si, _, _ = random_problem(2, 30, 40, 64, 8, 2, 1, 0.0, 0.0)
    model = make_model(params, 2, 1, 0.0)
    fd = {'X': torch.zeros((0, 2), dtype=torch.int64).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0}
    assert model.predict(fd)['prediction'].shape == (0,)


def test_confounders_come_from_the_torch_cpu_stream():
    """Without an injected 'sample_item' the model draws torch.randint(item_num, (P, S)) on the CPU generator,
    like src/models/DCCF.py:72 — same indices as the reference for the same seed and call sequence."""
    params, X, si, _, _ = random_problem(2, 30, 40, 64, 8, 4, 1, 0.0, 0.0)
    model = make_model(params, 4, 1, 0.0)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0}
    torch.manual_seed(123)
    a = model.predict(fd)['prediction'].cpu().numpy()
    torch.manual_seed(123)
    draw = torch.randint(40, size=(8, 4))
    fd2 = dict(fd)
    fd2['sample_item'] = draw
    b = model.predict(fd2)['prediction'].cpu().numpy()
    assert np.array_equal(a, b)
    want = O.MT19937(123).torch_randint(40, 32).reshape(8, 4)
    assert np.array_equal(draw.numpy(), want)


def test_graph_replay_equals_eager_steps():
    """The CUDA-graph replay of the fused step (device-resident step / rng counters) is bit-identical to
    launching the kernels one by one, over several steps and with an evaluation call in between."""
    U, I, F, P, S, A, std, drop = 200, 300, 768, 64, 10, 2, 0.1, 0.2
    params, X, si, _, _ = random_problem(8, U, I, F, P, S, A, 0.0, 0.0)
    rs = np.random.RandomState(3)
    outs = []
    for use_graph in (False, True):
        model = make_model(params, S, A, std)
        model.use_cuda_graph = use_graph
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        losses = []
        for t in range(6):
            Xt = np.roll(X, t, axis=0).copy()
            Xt[P // 2:, 0] = Xt[:P // 2, 0]
            fd = {'X': torch.from_numpy(Xt).cuda(), 'rank': 1, 'train': True, 'dropout': drop,
                  'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(np.roll(si, t, axis=0).copy())}
            losses.append(float(model.train_step(fd)['loss']))
            if t == 3:       # an evaluation pass advances the rng call counter between training steps
                model.predict({'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
                               'sample_item': torch.from_numpy(si)})
        outs.append((losses, model_params(model), model.optimizer.step_count))
    assert outs[0][0] == outs[1][0]
    assert outs[0][2] == outs[1][2] == 6
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k


# ---------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05, 3xTF32) evaluation scorer
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('P,F', [(300, 768), (37, 128), (6, 64)])
def test_tensor_core_scorer_vs_oracle(P, F):
    """dccf_score_fwd_tc with the explicit noise tensor: the raw accumulator equals noise·W_f^T to 3xTF32
    accuracy (~1e-6) and the prediction keeps the 1e-5 bound; ragged tiles included."""
    U, I, S, A, std = 200, 300, 10, 2, 0.1
    params, X, si, noise, _ = random_problem(31, U, I, F, P, S, A, std, 0.0)
    model = make_model(params, S, A, std)
    N = P * (S + 1) * A
    dbg = torch.zeros((N, 64), device='cuda')
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(si), 'noise': torch.from_numpy(noise), 'force_tc': True, 'dbg_pre': dbg}
    got = model.predict(fd)['prediction'].cpu().numpy()
    want_pre = noise.astype(np.float64) @ params['W'][:, 64:].astype(np.float64).T
    assert rel_err(dbg.cpu().numpy(), want_pre) < 3e-6
    ref = O.predict(params, X, si, noise, None, A, dtype=np.float64)
    assert rel_err(got, ref['pred']) < 1e-5
    model.check_ids()


def test_tensor_core_scorer_equals_simt_scorer_on_library_noise():
    """Noise generated inside the kernels (mode 2): the tcgen05 path and the FP32 SIMT path see the same
    Philox stream and agree within the parity bound; with dropout too; tables follow parameter updates."""
    U, I, F, P, S, A, std = 500, 800, 768, 1024, 10, 2, 0.1
    params, X, si, _, _ = random_problem(41, U, I, F, P, S, A, 0.0, 0.0)
    model = make_model(params, S, A, std)
    for drop in (0.0, 0.3):
        fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': drop,
              'sample_item': torch.from_numpy(si)}
        model.use_tensor_cores = False
        model._rng_offset = 10
        a = model.predict(fd)['prediction'].cpu().numpy()
        model.use_tensor_cores = True
        model._rng_offset = 10
        b = model.predict(dict(fd, force_tc=True))['prediction'].cpu().numpy()
        assert rel_err(b, a) < 1e-5
    # a training step changes W / E_item: the projected tables must be rebuilt
    model.optimizer = model.make_fused_optimizer(lr=1e-2, l2=1e-4)
    tr = {'X': torch.from_numpy(X[:64]).cuda(), 'rank': 1, 'train': True, 'dropout': 0.2, 'Y': torch.zeros(64).cuda(),
          'sample_item': torch.from_numpy(si[:64])}
    model.train_step(tr)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0, 'sample_item': torch.from_numpy(si)}
    model.use_tensor_cores = False
    model._rng_offset = 50
    a = model.predict(fd)['prediction'].cpu().numpy()
    model.use_tensor_cores = True
    model._rng_offset = 50
    b = model.predict(dict(fd, force_tc=True))['prediction'].cpu().numpy()
    assert rel_err(b, a) < 1e-5


@pytest.mark.parametrize('U,I,F,P,S,A', [(300, 500, 768, 256, 10, 2), (50, 70, 128, 38, 3, 1), (20, 30, 64, 2, 0, 1),
                                         (64, 64, 256, 130, 5, 3)])
def test_tensor_core_training_step_equals_simt_step(U, I, F, P, S, A):
    """Three fused training steps (library noise + dropout, mode 2) with the contractions on the tensor cores
    (dccf_train_fwd_tc / dccf_train_bwd_tc) against the same steps on the FP32 SIMT kernels: same Philox streams,
    predictions / loss / Adam moments / updated tables within the parity bounds."""
    params, X, si, _, _ = random_problem(23, U, I, F, P, S, A, 0.0, 0.0)
    X[P // 2:, 0] = X[:P // 2, 0]
    outs = []
    for tc in (False, True):
        model = make_model(params, S, A, 0.1)
        model.use_tensor_cores_train = tc
        model.use_cuda_graph = False
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        preds, losses = [], []
        for t in range(3):
            fd = {'X': torch.from_numpy(np.roll(X, t, axis=0).copy()).cuda(), 'rank': 1, 'train': True, 'dropout': 0.2,
                  'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(np.roll(si, t, axis=0).copy())}
            fd['X'][P // 2:, 0] = fd['X'][:P // 2, 0]
            o = model.train_step(fd)
            preds.append(o['prediction'].cpu().numpy().copy())
            losses.append(float(o['loss']))
        opt = model.optimizer
        outs.append((preds, losses, model_params(model), {k: v.cpu().numpy() for k, v in opt.exp_avg.items()},
                     {k: v.cpu().numpy() for k, v in opt.exp_avg_sq.items()}))
        model.check_ids()
    a, b = outs
    for t in range(3):
        assert rel_err(b[0][t], a[0][t]) < 1e-5
        assert abs(b[1][t] - a[1][t]) < 1e-5 * abs(a[1][t])
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert rel_err(b[3][k], a[3][k]) < 2e-5, k          # exp_avg
        assert rel_err(b[4][k], a[4][k]) < 2e-5, k          # exp_avg_sq
    for k in ('E_user', 'E_item'):
        assert rel_err(b[2][k], a[2][k]) < 1e-5, k
    assert rel_err(b[2]['W'], a[2]['W']) < 5e-4              # see DESIGN.md §4 (lr/eps sensitivity where |g| << eps)


# ---------------------------------------------------------------------------------------------------------
# full-catalogue scoring (tcgen05 GEMM + fused top-k)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('U,I', [(300, 1000), (129, 67), (1000, 4096)])
def test_full_catalogue_ipsmf_scores_and_topk(U, I):
    """IPSBiasedMF over the whole catalogue (src/models/IPSBiasedMF.py:37-57): matrix vs float64 numpy,
    fused top-k ids bit-identical to the top-k of the materialised matrix (ties by item id)."""
    from dccf_b200 import full_catalogue, synth
    fac = synth.make_ipsmf_factors(U, I, seed=11)
    dev = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else float(v)) for k, v in fac.items()}
    got = full_catalogue.ipsmf_exposure(dev)
    want = O.exposure_values({k: (v.astype(np.float64) if isinstance(v, np.ndarray) else float(v)) for k, v in fac.items()},
                             np.arange(U), np.tile(np.arange(I), (U, 1)))
    assert got.shape == (U, I)
    assert rel_err(got.cpu().numpy(), want) < 3e-6
    for k in (1, 5, 16):
        kk = min(k, I)
        s, ids = full_catalogue.ipsmf_topk(dev, kk)
        m = got.cpu().numpy()
        order = np.lexsort((np.tile(np.arange(I), (U, 1)), -m.astype(np.float64)), axis=1)[:, :kk]
        assert np.array_equal(ids.cpu().numpy(), order)
        assert np.array_equal(s.cpu().numpy(), np.take_along_axis(m, order, axis=1))


def test_full_catalogue_dccf_topk_matches_pairwise_scorer():
    """Deterministic DCCF (std 0, no confounders): the GEMM scorer ranks the catalogue like the pairwise scorer."""
    from dccf_b200 import full_catalogue
    U, I, F = 64, 500, 768
    params, _, _, _, _ = random_problem(5, U, I, F, 2, 0, 1, 0.0, 0.0)
    model = make_model(params, 0, 1, 0.0)
    s, ids = full_catalogue.dccf_catalogue_topk(model, 5)
    X = np.stack([np.repeat(np.arange(U), I), np.tile(np.arange(I), U)], 1).astype(np.int64)
    pair = model.predict({'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0})
    pair = pair['prediction'].cpu().numpy().reshape(U, I)
    top_pair = np.take_along_axis(pair, ids.cpu().numpy(), axis=1)
    assert rel_err(s.cpu().numpy(), top_pair) < 1e-5
    # the 5th best pairwise score is not beaten by anything outside the returned set (up to rounding)
    kth = np.sort(pair, axis=1)[:, -5]
    assert (top_pair.min(axis=1) >= kth - 1e-5 * np.abs(pair).max()).all()


def test_predict_many_equals_sequential_predict():
    """The evaluation loop with the prefetching worker thread consumes the torch CPU generator in the same
    order as batch-by-batch predict calls (src/models/DCCF.py:72): identical confounders, identical scores."""
    U, I, F, S, A, std = 100, 150, 128, 10, 2, 0.1
    params, X, _, _, _ = random_problem(13, U, I, F, 300, S, A, 0.0, 0.0)
    bounds = [(0, 128), (128, 256), (256, 300)]
    outs = []
    for many in (False, True):
        model = make_model(params, S, A, std)
        torch.manual_seed(99)
        fds = [{'X': torch.from_numpy(X[a:b]).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0} for a, b in bounds]
        if many:
            preds = model.predict_many(fds)
        else:
            preds = [model.predict(fd)['prediction'] for fd in fds]
        outs.append(torch.cat(preds).cpu().numpy())
        after = torch.randint(1000, (4,))
        outs.append(after.numpy())
    assert np.array_equal(outs[0], outs[2]) and np.array_equal(outs[1], outs[3])


def test_resident_epoch_equals_step_by_step():
    """The epoch-resident path (device-side batch cursor, one graph launch per step, one confounder draw for
    the whole epoch) is bit-identical to feeding the same batches one by one."""
    U, I, F, P, S, A, std, drop = 200, 300, 768, 64, 10, 2, 0.1, 0.2
    params, X, _, _, _ = random_problem(8, U, I, F, P, S, A, 0.0, 0.0)
    n = 5
    Xs = np.stack([np.roll(X, t, axis=0) for t in range(n)])
    Xs[:, P // 2:, 0] = Xs[:, :P // 2, 0]
    outs = []
    for resident in (False, True):
        model = make_model(params, S, A, std)
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        torch.manual_seed(4242)
        losses = []
        if resident:
            draws = model.draw_confounders(n * P).view(n, P, S)
            step = model.begin_resident_epoch(torch.from_numpy(Xs).cuda(), draws.cuda(), drop)
            losses.append(float(step.first['loss']))
            while step.remaining() > 0:
                losses.append(float(step()['loss']))
        else:
            for t in range(n):
                fd = {'X': torch.from_numpy(Xs[t]).cuda(), 'rank': 1, 'train': True, 'dropout': drop,
                      'Y': torch.zeros(P).cuda()}
                losses.append(float(model.train_step(fd)['loss']))
        outs.append((losses, model_params(model), torch.randint(1000, (3,)).numpy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][2], outs[1][2])           # the CPU generator ends in the same state
    for k in ('E_user', 'E_item', 'W', 'b'):
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), k


def test_fused_adam_step_record_lists():
    """dccf_adam_step (the kernel the training step uses): rows with 1, 2, 50 (ranked in shared memory) and 300
    (longer than the list buffer: repeated walks) records, two tables + two dense tensors in one launch; result
    identical to the per-tensor entry points and to the oracle."""
    from dccf_b200 import kernels
    rs = np.random.RandomState(12)
    rows, n_rec = 700, 900
    keys = rs.randint(0, rows, size=n_rec).astype(np.int32)
    keys[:300] = 17
    keys[300:350] = 99
    keys = keys[rs.permutation(n_rec)]
    grads = (rs.standard_normal((n_rec, 64)) * 0.01).astype(np.float32)
    table = (rs.standard_normal((rows, 64)) * 0.05).astype(np.float32)
    dense = np.zeros((rows, 64), dtype=np.float32)
    for r in range(n_rec):
        dense[keys[r]] += grads[r]
    want_p, want_m, want_v = O.adam_step(table, dense, np.zeros_like(table), np.zeros_like(table), 3)
    Wd = (rs.standard_normal(5000) * 0.05).astype(np.float32)
    gparts = (rs.standard_normal((4, 5000)) * 0.01).astype(np.float32)
    want_W, _, _ = O.adam_step(Wd, gparts.sum(0), np.zeros_like(Wd), np.zeros_like(Wd), 3)
    hp = kernels.make_adam(1e-3, 1e-4, 1e-4, step=3)
    res = []
    for fused in (True, False):
        t = torch.from_numpy(table).cuda(); m = torch.zeros_like(t); v = torch.zeros_like(t)
        w = torch.from_numpy(Wd).cuda(); wm = torch.zeros_like(w); wv = torch.zeros_like(w)
        head = torch.full((rows,), -1, dtype=torch.int32, device='cuda')
        nxt = torch.empty(n_rec, dtype=torch.int32, device='cuda')
        k_d, g_d, gp_d = torch.from_numpy(keys).cuda(), torch.from_numpy(grads).cuda(), torch.from_numpy(gparts).cuda()
        if fused:
            kernels.adam_step([kernels.adam_table(t, m, v, k_d, g_d, 1, n_rec, n_rec, n_rec * 64, head, nxt)],
                              [kernels.adam_tensor(w, wm, wv, gp_d, 4, 5000)], hp)
        else:
            kernels.adam_sweep(t, m, v, k_d, g_d, n_rec, head, nxt, hp)
            kernels.adam_dense(w, wm, wv, gp_d, 4, 5000, hp)
        assert int((head != -1).sum()) == 0
        res.append((t.cpu().numpy(), m.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy()))
    for a, b in zip(res[0], res[1]):
        assert np.array_equal(a, b)
    assert rel_err(res[0][1], want_m) < 1e-6 and rel_err(res[0][2], want_v) < 1e-6
    assert rel_err(res[0][0], want_p) < 1e-5 and rel_err(res[0][3], want_W) < 1e-5


def test_csr_record_lists_equal_linked_lists_with_hot_rows():
    """The split optimizer step with CSR record lists (counting link + dccf_adam_csr_build) against the same step with
    linked lists: rows with 1, a few, ~100 and > 700 records (the hot item of every slot of 64 pairs — beyond the
    shared-memory ranking buffer), eager and CUDA-graph replay: identical bits (same summation order)."""
    U, I, F, P, S, A = 50, 60, 128, 128, 10, 2
    params, X, si, _, _ = random_problem(31, U, I, F, P, S, A, 0.1, 0.2)
    si[:64, :] = 7                          # 640 confounder slots on item 7
    X[:64, 1] = 7                           # + 64 true-item slots
    si[64:, :3] = 11                        # ~190 records on item 11
    X[:10, 0] = 3                           # a user with many pairs
    X[64:74, 0] = 3
    outs = []
    for csr in (True, False):
        model = make_model(params, S, A, 0.1)
        model.use_csr_lists = csr
        model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
        for t in range(4):                  # step 1 eager, step 2 captures, steps 3-4 replay the graph
            fd = {'X': torch.from_numpy(np.roll(X, t, axis=0).copy()).cuda(), 'rank': 1, 'train': True, 'dropout': 0.2,
                  'Y': torch.zeros(P).cuda(), 'sample_item': torch.from_numpy(np.roll(si, t, axis=0).copy())}
            out = model.train_step(fd)
        torch.cuda.synchronize()
        model.check_ids()
        assert int((model.optimizer.head_i != -1).sum()) == 0 and int((model.optimizer.head_u != -1).sum()) == 0
        if csr:
            assert int(model.optimizer.csr_pool.abs().sum()) == 0          # the ranges were released
        outs.append((model_params(model), float(out['loss'])))
    for k in outs[0][0]:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), k
    assert outs[0][1] == outs[1][1]


def test_fused_training_follows_the_reference_at_config0(tmp_path):
    """BASELINE.json configs[0] at full size on the GPU: the first 24 training steps of the UNMODIFIED reference
    (tests/golden/config0_train.npz) replayed through model.train_step — 256 pairs = 5 632 predictor rows x 832 inputs
    per step, Adam over 2 000 + 5 000 embedding rows — from seeds alone: dataset, initial weights, batches and
    negatives from the host pipeline, confounders / noise / dropout masks re-drawn from the torch CPU generator call
    for call.  Loss of every step within 1e-5; final parameters within the drift the float64 oracle itself shows
    against the fp32 reference after 24 Adam steps (tests/test_oracle_golden.py::test_oracle_follows_the_reference_at_config0)."""
    from conftest import config0_draws, config0_problem
    steps = 24
    g, model, feat, expo, Xs = config0_problem(str(tmp_path), steps)
    if str(g['torch_version']) != torch.__version__:
        pytest.skip('random inputs are re-drawn from the torch CPU generator: needs torch %s' % g['torch_version'])
    assert np.array_equal(Xs[0], g['X_first']) and np.array_equal(Xs[-1], g['X_last'])
    model = model.cuda()
    model.optimizer = model.make_fused_optimizer(lr=float(g['lr']), l2=float(g['l2']))
    drop = float(g['dropout'])
    b = Xs[0].shape[0] // 2
    Y = torch.cat([torch.ones(b), torch.zeros(b)]).cuda()
    for t, (si, noise, mask) in enumerate(config0_draws(g, model.item_num, steps)):
        out = model.train_step({'X': torch.from_numpy(Xs[t]).cuda(), 'Y': Y, 'rank': 1, 'train': True, 'dropout': drop,
                                'sample_item': si, 'noise': noise, 'dropout_mask': mask})
        assert abs(float(out['loss']) - g['loss'][t]) < 1e-5 * abs(g['loss'][t]), t
        if t == 0:
            assert rel_err(out['prediction'].cpu().numpy(), g['pred_first']) < 1e-5
    assert rel_err(out['prediction'].cpu().numpy(), g['pred_last']) < 2e-4
    got = model_params(model)
    assert rel_err(got['W'][::8], g['final_W_rows']) < 2e-4
    assert rel_err(got['E_user'][g['users_last']], g['final_E_user_rows']) < 4e-4
    assert rel_err(got['E_item'][g['items_last']], g['final_E_item_rows']) < 4e-4
    assert rel_err(got['b'], g['final_b']) < 1e-3
    norms = [np.linalg.norm(got[k].astype(np.float64)) for k in ('E_user', 'E_item', 'W', 'b')]
    assert np.abs(np.array(norms) / g['norms'] - 1).max() < 1e-5
    model.check_ids()


def test_evaluation_follows_the_reference_at_config0(tmp_path):
    """BASELINE.json configs[0], evaluation on the GPU: the reference's own predictions and ndcg@5 / recall@5 /
    precision@5 for 100 test users x (positives + 1000 negatives) (tests/golden/config0_eval.npz), reproduced through
    model.predict (full 16 384-pair batches on the tcgen05 scorer, explicit noise re-drawn from the torch CPU generator
    call for call) and the device ranker: predictions within 1e-5, metrics within 1e-6."""
    from conftest import config0_eval_draws, config0_eval_problem
    from dccf_b200.models.BaseModel import BaseModel
    g, model, feat, expo, sub = config0_eval_problem(str(tmp_path))
    if str(g['torch_version']) != torch.__version__:
        pytest.skip('random inputs are re-drawn from the torch CPU generator: needs torch %s' % g['torch_version'])
    model = model.cuda()
    model.eval()
    preds = []
    for a, b, si, noise in config0_eval_draws(g, model.item_num, len(sub['Y'])):
        out = model.predict({'X': torch.from_numpy(sub['X'][a:b]).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
                             'sample_item': si, 'noise': noise})
        preds.append(out['prediction'])
    pred = torch.cat(preds)
    model.check_ids()
    assert rel_err(pred.cpu().numpy(), g['pred']) < 1e-5
    vals = BaseModel.evaluate_method(pred, sub, [str(m) for m in g['metrics']])
    assert np.abs(np.array(vals) - g['values']).max() < 1e-6


def test_ipsmf_exposure_matches_reference(golden):
    """The reference's IPSBiasedMF.predict over a whole U x I grid (tests/golden/ipsmf.npz, src/models/IPSBiasedMF.py:
    37-57) against the full-catalogue tcgen05 GEMM (the ips_expo_prob.npy producer) and against the on-the-fly
    exposure of the pairwise scorer (exposure mode 1): a DCCF model fed those factors equals a DCCF model fed the
    reference's dense matrix."""
    from dccf_b200 import full_catalogue
    g = golden('ipsmf')
    fac = {k: (g[k] if g[k].ndim else float(g[k])) for k in ('mf_user', 'mf_item', 'mf_user_bias', 'mf_item_bias',
                                                              'mf_global_bias', 'propensity', 'mf_min_propensity')}
    dev = {k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if isinstance(v, np.ndarray) else v) for k, v in fac.items()}
    got = full_catalogue.ipsmf_exposure(dev)
    assert rel_err(got.cpu().numpy(), g['pred']) < 3e-6
    U, I = g['pred'].shape
    F, P, S, A = 64, 48, 10, 2
    params, X, si, noise, _ = random_problem(5, U, I, F, P, S, A, 0.1, 0.0)
    fd = {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
          'sample_item': torch.from_numpy(si), 'noise': torch.from_numpy(noise)}
    a = make_model(params, S, A, 0.1, expo_factors=fac).predict(fd)['prediction'].cpu().numpy()
    b = make_model(dict(params, expo=g['pred']), S, A, 0.1).predict(fd)['prediction'].cpu().numpy()
    # exposure values reach |29| here (propensity clamp 0.1): an fp32 ulp of the logit moves a softmax weight by ~3e-6
    assert rel_err(a, b) < 3e-5


def test_whole_dccf_run_equals_reference_run(golden, tmp_path):
    """src/main.py's whole sequence with DCCF on the GPU == the same sequence of the UNMODIFIED reference in the
    deterministic configuration --std 0 --dropout 0 (tests/golden/run_dccf.npz, oracle/make_golden.py::
    make_run_fixture_dccf): same shuffles and negatives (numpy generator), same confounder draws (torch CPU generator,
    consumed in the same order by evaluation passes and training steps), hence the same "Test Before Training"
    ndcg@5 / recall@5 / precision@5 to 1e-6, the same per-epoch train rmse / mae, and validation / test metrics,
    final predictions and checkpoint within the drift of 28 fp32 Adam steps (a near-tie in a 21-candidate ranking may
    flip: 1/600 per flip)."""
    from dccf_b200 import synth
    from dccf_b200.data_loaders.DataLoader import DataLoader
    from dccf_b200.data_processor.DataProcessor import DataProcessor
    from dccf_b200.models.DCCF import DCCF
    from dccf_b200.runners.BaseRunner import BaseRunner
    g = golden('run_dccf')
    seed = int(g['seed'])
    d = synth.write_dataset(str(tmp_path), 'toy', int(g['n_users']), int(g['n_items']), int(g['per_user']), feat_dim=64,
                            seed=seed + 5)
    torch.manual_seed(seed)
    np.random.seed(seed)
    dl = DataLoader(path=str(tmp_path), dataset='toy', label='label', sep=',')
    model = DCCF(path=d, dataset='toy', sentence_model=synth.DEFAULT_SENTENCE_MODEL, sample_num=10, attribute_num=2,
                 std=0.0, label_min=dl.label_min, label_max=dl.label_max, feature_num=0, user_num=dl.user_num,
                 item_num=dl.item_num, u_vector_size=64, i_vector_size=64, n_layers=1, random_seed=seed,
                 model_path=str(tmp_path / 'model' / 'm.pt'))
    model.apply(model.init_paras)
    model = model.cuda()
    dl.drop_neg()
    dp = DataProcessor(dl, model, rank=1, test_neg_n=int(g['test_neg_n']))
    runner = BaseRunner(optimizer='Adam', learning_rate=float(g['lr']), epoch=int(g['epochs']),
                        batch_size=int(g['batch_size']), eval_batch_size=16384, dropout=0.0, l2=float(g['l2']),
                        metrics='ndcg@5,recall@5,precision@5', check_epoch=1, early_stop=1)
    runner.show_progress = False
    before = runner.evaluate(model, dp.get_test_data(), dp)
    runner.train(model, dp, skip_eval=0)
    after = runner.evaluate(model, dp.get_test_data(), dp)
    pred = runner.predict(model, dp.get_test_data(), dp)
    model.check_ids()
    assert np.abs(np.array(before) - g['before']).max() < 1e-6
    assert np.asarray(runner.train_results).shape == g['train_results'].shape
    assert np.abs(np.array(runner.train_results) - g['train_results']).max() < 1e-4
    assert np.abs(np.array(runner.valid_results) - g['valid_results']).max() < 0.01
    assert np.abs(np.array(runner.test_results) - g['test_results']).max() < 0.01
    assert np.abs(np.array(after) - g['after']).max() < 0.01
    assert rel_err(pred, g['pred']) < 2e-3
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    assert sorted('sd_' + k for k in sd) == sorted(k for k in g.files if k.startswith('sd_'))
    for k, v in sd.items():
        assert rel_err(v, g['sd_' + k]) < 2e-3, k
