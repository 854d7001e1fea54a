"""The numpy oracle (oracle/dccf_oracle.py) pinned against fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, written by oracle/make_golden.py) and against the reference's docstring values."""
import numpy as np
import pytest

from conftest import golden_params, rel_err
from oracle import dccf_oracle as O

TRAIN_FIXTURES = ['train_f64', 'train_f768', 'train_nodrop']


@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_predict_matches_reference(golden, name):
    g = golden(name)
    A, drop = int(g['A']), float(g['dropout'])
    params = golden_params(g)
    mask = g['mask_0'] if drop > 0 else None
    out = O.predict(params, g['X_0'], g['sample_item_0'], g['noise_0'], mask, A)
    assert rel_err(out['pred'], g['pred_0']) < 1e-5                     # north_star: 1e-5 relative, fp32
    assert abs(O.loss_bpr(out['pred']) - float(g['loss_0'])) < 1e-5 * abs(float(g['loss_0']))


@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_eval_predict_matches_reference(golden, name):
    g = golden(name)
    params = golden_params(g, 'final_')
    out = O.predict(params, g['eval_X'], g['eval_sample_item'], g['eval_noise'], None, int(g['A']))
    assert rel_err(out['pred'], g['eval_pred']) < 1e-5


@pytest.mark.parametrize('name', TRAIN_FIXTURES)
def test_training_steps_match_reference(golden, name):
    """Forward, backward, l2 + clip + Adam restated; compared after `steps` reference steps."""
    g = golden(name)
    A, steps, drop = int(g['A']), int(g['steps']), float(g['dropout'])
    params = golden_params(g)
    state = {k: {'m': np.zeros_like(params[k]), 'v': np.zeros_like(params[k])} for k in ('E_user', 'E_item', 'W', 'b')}
    hp = dict(lr=float(g['lr']), l2=float(g['l2']), weight_decay=float(g['l2']))
    for t in range(steps):
        mask = g['mask_%d' % t] if drop > 0 else None
        params, state, loss, pred = O.train_step(params, state, t + 1, g['X_%d' % t], g['sample_item_%d' % t],
                                                 g['noise_%d' % t], mask, A, hp)
        assert rel_err(pred, g['pred_%d' % t]) < 2e-5
        assert abs(loss - float(g['loss_%d' % t])) < 1e-5 * abs(float(g['loss_%d' % t]))
    for k in ('E_user', 'E_item', 'W', 'b'):
        # exp_avg / exp_avg_sq are linear / quadratic in the gradient: well conditioned
        assert rel_err(state[k]['m'], g['m_' + k]) < 2e-5, k
        assert rel_err(state[k]['v'], g['v_' + k]) < 2e-5, k
    # the parameters: p -= lr * m_hat / (sqrt(v_hat) + 1e-8).  Where |g| << eps (elements of W whose data
    # gradient cancels the l2 term) d(update)/dg = lr/eps = 1e5, so fp32 summation noise of 1e-10 in g moves
    # the weight by 1e-5 absolute: the reference differs from itself across BLAS builds by that much.
    assert rel_err(params['E_user'], g['final_E_user']) < 1e-5
    assert rel_err(params['E_item'], g['final_E_item']) < 1e-5
    assert rel_err(params['b'], g['final_b']) < 1e-5
    assert rel_err(params['W'], g['final_W']) < 5e-4


def test_ranking_metrics_match_reference(golden):
    g = golden('metrics')
    data = {'uid': g['uid'], 'iid': g['iid'], 'Y': g['Y']}
    vals = O.evaluate_method(g['p'], data, [str(m) for m in g['metrics']])
    assert np.abs(np.array(vals) - g['values']).max() < 1e-6             # north_star: metrics to 1e-6


def test_metric_known_answers(golden):
    """Values printed in the reference's docstrings (src/utils/rank_metrics.py:64-70,143-144,179-187)."""
    g = golden('metrics')
    assert abs(O.ndcg_at_k([2, 1, 2, 0], 4) - 0.96519546960144276) < 1e-12
    assert abs(O.dcg_at_k([3, 2, 3, 0, 0, 1, 2, 2, 3, 0], 2) - 4.2618595071429155) < 1e-12
    assert O.ndcg_at_k([0], 1) == 0.0
    assert O.ndcg_at_k([1], 2) == 1.0
    assert abs(float(g['kat_ndcg_2120_k4_m1']) - 0.96519546960144276) < 1e-12
    assert abs(float(g['kat_dcg_k2_m1']) - 4.2618595071429155) < 1e-12
    assert abs(float(g['kat_prec_001_k3']) - 1.0 / 3.0) < 1e-12


def test_rank_users_tie_break():
    """Total order: score desc, then item id asc, then row asc; NaN last."""
    scores = np.array([0.5, 0.5, 0.5, np.nan, 0.9, 0.5], dtype=np.float32)
    uid = np.zeros(6, dtype=np.int64)
    iid = np.array([7, 3, 3, 1, 9, 2])
    Y = np.array([0, 1, 0, 1, 0, 0], dtype=np.float32)
    users, topk, rows, m = O.rank_users(scores, uid, Y, iid, 6)
    assert list(topk[0]) == [9, 2, 3, 3, 7, 1]
    assert list(rows[0]) == [4, 5, 1, 2, 0, 3]


def test_mt19937_matches_numpy_and_torch():
    """SURVEY.md Appendix C recipes: numpy legacy randint / torch CPU randint from raw MT19937 words."""
    import torch
    np.random.seed(2019)
    mt = O.MT19937(2019)
    assert [int(np.random.randint(5000)) for _ in range(50)] == [mt.np_randint(5000) for _ in range(50)]
    torch.manual_seed(77)
    want = torch.randint(16000, size=(4, 10)).reshape(-1).numpy()
    got = O.MT19937(77).torch_randint(16000, 40)
    assert np.array_equal(want, got)


@pytest.mark.parametrize('name', ['train_f64', 'train_nodrop'])
def test_torch_port_matches_reference(golden, name):
    """oracle/torch_port.py (bench.py's CPU baseline arm) reproduces the reference's training trajectory."""
    import torch
    from oracle import torch_port
    g = golden(name)
    A, S, steps, drop = int(g['A']), int(g['S']), int(g['steps']), float(g['dropout'])
    U, I = g['init_E_user'].shape[0], g['init_E_item'].shape[0]
    model = torch_port.DCCFPort(U, I, g['feat'], g['expo'], sample_num=S, attribute_num=A, std=float(g['std']),
                                seed=int(g['seed']))
    assert np.array_equal(model.mlp[0].weight.detach().numpy(), g['init_W'])       # same init stream
    optim = torch.optim.Adam(model.parameters(), lr=float(g['lr']), weight_decay=float(g['l2']))
    for t in range(steps):
        X = torch.from_numpy(g['X_%d' % t])
        fd = {'X': X, 'Y': torch.zeros(X.shape[0]), 'dropout': drop,
              'sample_item': torch.from_numpy(g['sample_item_%d' % t]), 'noise': torch.from_numpy(g['noise_%d' % t])}
        if drop > 0:
            fd['dropout_mask'] = torch.from_numpy(g['mask_%d' % t])
        out = torch_port.fit_step(model, optim, fd, float(g['l2']))
        assert rel_err(out['prediction'].detach().numpy(), g['pred_%d' % t]) < 1e-6
    assert rel_err(model.uid_embeddings.weight.detach().numpy(), g['final_E_user']) < 1e-6
    assert rel_err(model.mlp[0].weight.detach().numpy(), g['final_W']) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# property tests of the oracle's ranker (the checker of the device ranker): SURVEY.md §4 plan item (4)
# ---------------------------------------------------------------------------------------------------------
def _brute_force_user(scores, labels, iids, rows, k):
    """One user, straight from the definition: candidates in the total order (score desc, item id asc, row asc),
    NaN scores last; ndcg@k with method-1 DCG and the ideal from the user's own labels
    (src/utils/rank_metrics.py:159-164,198), hit / precision / recall / f1 as src/models/BaseModel.py:99-126."""
    import functools
    import math

    def cmp(a, b):
        sa, sb = scores[a], scores[b]
        na, nb = math.isnan(sa), math.isnan(sb)
        if na != nb:
            return 1 if na else -1
        if not na and sa != sb:
            return -1 if sa > sb else 1
        if iids[a] != iids[b]:
            return -1 if iids[a] < iids[b] else 1
        return -1 if rows[a] < rows[b] else 1

    order = sorted(range(len(scores)), key=functools.cmp_to_key(cmp))
    rel = [float(labels[i]) for i in order]
    dcg = sum(r / math.log2(j + 2) for j, r in enumerate(rel[:k]))
    ideal = sum(r / math.log2(j + 2) for j, r in enumerate(sorted(rel, reverse=True)[:k]))
    hits = sum(rel[:k])
    tot = sum(rel)
    return ([iids[i] for i in order[:k]], [rows[i] for i in order[:k]],
            [dcg / ideal if ideal else 0.0, 1.0 if hits > 0 else 0.0, sum(1 for r in rel[:k] if r != 0) / k,
             hits / tot if tot else float('nan'), 2.0 * hits / (k + tot)])


def test_rank_users_property_against_brute_force():
    from hypothesis import given, settings, strategies as st

    # few distinct scores and item ids: ties everywhere; NaN allowed; users of very different sizes
    cand = st.tuples(st.integers(0, 3), st.sampled_from([0.0, 0.5, 0.5000001, -1.0, 2.0, float('nan')]),
                     st.integers(0, 5), st.booleans())

    @settings(max_examples=200, deadline=None)
    @given(st.lists(cand, min_size=1, max_size=40), st.integers(1, 7))
    def check(cands, k):
        uid = np.array([c[0] for c in cands], dtype=np.int64)
        scores = np.array([c[1] for c in cands], dtype=np.float32)
        iid = np.array([c[2] for c in cands], dtype=np.int64)
        Y = np.array([1.0 if c[3] else 0.0 for c in cands], dtype=np.float32)
        users, topk, rows, m = O.rank_users(scores, uid, Y, iid, k)
        assert list(users) == sorted(set(uid.tolist()))
        for g, u in enumerate(users):
            mine = [i for i in range(len(cands)) if uid[i] == u]
            want_iid, want_row, want_m = _brute_force_user([float(scores[i]) for i in mine], [Y[i] for i in mine],
                                                           [int(iid[i]) for i in mine], mine, k)
            n = len(want_iid)
            assert list(topk[g, :n]) == want_iid and all(topk[g, n:] == -1)
            assert list(rows[g, :n]) == want_row
            for a, b in zip(m[g], want_m):
                assert (np.isnan(a) and np.isnan(b)) or abs(a - b) < 1e-12

    check()


def test_oracle_follows_the_reference_at_config0(tmp_path):
    """BASELINE.json configs[0] at full size (2 000 users x 5 000 items x 768-d, 256 pairs = 5 632 predictor rows per
    step): 24 training steps of the UNMODIFIED reference (tests/golden/config0_train.npz, oracle/make_golden.py::
    make_config0_train) replayed by the oracle from seeds alone — dataset, initial weights, batches and negatives (host
    pipeline), confounders, noise and dropout masks (the torch CPU generator re-drawn call for call).  Per step the loss
    agrees to 1e-6; parameters drift apart at Adam's lr/eps conditioning (float64 gradients here, fp32 there)."""
    import torch
    from conftest import config0_draws, config0_problem
    steps = 24
    g, model, feat, expo, Xs = config0_problem(str(tmp_path), steps)
    if str(g['torch_version']) != torch.__version__:
        pytest.skip('random inputs are re-drawn from the torch CPU generator: needs torch %s' % g['torch_version'])
    assert np.array_equal(Xs[0], g['X_first']) and np.array_equal(Xs[-1], g['X_last'])     # batches + negatives
    params = {'E_user': model.uid_embeddings.weight.detach().numpy().copy(),
              'E_item': model.iid_embeddings.weight.detach().numpy().copy(),
              'W': model.mlp[0].weight.detach().numpy().copy(), 'b': model.mlp[0].bias.detach().numpy().copy(),
              'Feat': feat, 'expo': expo}
    state = {k: {'m': np.zeros_like(params[k]), 'v': np.zeros_like(params[k])} for k in ('E_user', 'E_item', 'W', 'b')}
    hp = dict(lr=float(g['lr']), l2=float(g['l2']), weight_decay=float(g['l2']))
    for t, (si, noise, mask) in enumerate(config0_draws(g, params['E_item'].shape[0], steps)):
        if t == 0:
            assert np.array_equal(si.numpy(), g['sample_item_first'])                       # confounders bit-exact
        p2, state, loss, pred = O.train_step(params, state, t + 1, Xs[t], si.numpy(), noise.numpy(), mask.numpy(),
                                             int(g['A']), hp)
        params = dict(p2, Feat=feat, expo=expo)
        assert abs(loss - g['loss'][t]) < 1e-6 * abs(g['loss'][t]), t
        if t == 0:
            assert rel_err(pred, g['pred_first']) < 1e-5
    assert rel_err(pred, g['pred_last']) < 2e-4
    assert rel_err(params['W'][::8], g['final_W_rows']) < 1e-4
    assert rel_err(params['E_user'][g['users_last']], g['final_E_user_rows']) < 2e-4
    assert rel_err(params['E_item'][g['items_last']], g['final_E_item_rows']) < 2e-4
    assert rel_err(params['b'], g['final_b']) < 5e-4
    norms = [np.linalg.norm(params[k].astype(np.float64)) for k in ('E_user', 'E_item', 'W', 'b')]
    assert np.abs(np.array(norms) / g['norms'] - 1).max() < 1e-5


def test_oracle_follows_the_reference_evaluation_at_config0(tmp_path):
    """BASELINE.json configs[0], evaluation: 100 test users x (positives + 1000 negatives) scored by the UNMODIFIED
    reference in batches of 16 384 pairs (tests/golden/config0_eval.npz, oracle/make_golden.py::make_config0_eval).
    The oracle ranker reproduces the reference's ndcg@5 / recall@5 / precision@5 from the reference's predictions, and
    the oracle scorer reproduces the predictions of the first two batches (32 768 pairs = 720 896 predictor rows) from
    seeds alone — ids from the host pipeline, confounders and noise re-drawn from the torch CPU generator."""
    import hashlib
    import torch
    from conftest import config0_eval_draws, config0_eval_problem
    g, model, feat, expo, sub = config0_eval_problem(str(tmp_path))
    if str(g['torch_version']) != torch.__version__:
        pytest.skip('random inputs are re-drawn from the torch CPU generator: needs torch %s' % g['torch_version'])
    assert len(sub['Y']) == int(g['n_rows'])
    X = np.ascontiguousarray(sub['X'])
    h = hashlib.sha256()
    h.update(str(X.dtype).encode() + b'|' + str(X.shape).encode() + b'|')
    h.update(X.tobytes())
    assert h.hexdigest() == str(g['X_digest'])                                   # the evaluation set itself
    vals = O.evaluate_method(g['pred'], sub, [str(m) for m in g['metrics']])
    assert np.abs(np.array(vals) - g['values']).max() < 1e-6                     # ranking metrics to 1e-6
    params = {'E_user': model.uid_embeddings.weight.detach().numpy(), 'E_item': model.iid_embeddings.weight.detach().numpy(),
              'W': model.mlp[0].weight.detach().numpy(), 'b': model.mlp[0].bias.detach().numpy(), 'Feat': feat,
              'expo': expo}
    for k, (a, b, si, noise) in enumerate(config0_eval_draws(g, params['E_item'].shape[0], len(sub['Y']))):
        pred = O.predict(params, sub['X'][a:b], si.numpy(), noise.numpy(), None, int(g['A']))['pred']
        assert rel_err(pred, g['pred'][a:b]) < 1e-5
        if k == 1:
            break


def test_ipsmf_exposure_matches_reference(golden):
    """oracle.exposure_values (dict form) == IPSBiasedMF.predict of the unmodified reference over a whole U x I grid
    (src/models/IPSBiasedMF.py:37-57), including propensities below, at and above the clamp M."""
    g = golden('ipsmf')
    fac = {k: (g[k] if g[k].ndim else float(g[k])) for k in ('mf_user', 'mf_item', 'mf_user_bias', 'mf_item_bias',
                                                              'mf_global_bias', 'propensity', 'mf_min_propensity')}
    U, I = g['pred'].shape
    got = O.exposure_values(fac, np.arange(U), np.tile(np.arange(I), (U, 1)))
    assert rel_err(got, g['pred']) < 1e-6
    fac64 = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in fac.items()}
    assert rel_err(O.exposure_values(fac64, np.arange(U), np.tile(np.arange(I), (U, 1))), g['pred']) < 1e-6
