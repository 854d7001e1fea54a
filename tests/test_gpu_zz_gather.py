"""Parity of the two inference paths added after round 1's GPU work: the noise-free gather scorer (dccf_score_gather,
csrc/gather_scores.cu) against the reference fixture, the oracle and the general FP32 scorer; and evaluation noise
drawn in the 64-d image of W_f (DCCF.eval_noise = 'projected') against the reference formula on the equivalent 768-d
noise tensor; and torch's CPU generator continued on the device (k_confounder_draw, host_rng.DeviceStream,
DCCF.device_confounders) against torch.randint itself.  Needs a GPU.

All cases were seen green on a B200 by the round-1 driver run; `DCCF.use_gather_scorer` is the default for noise-free
inference since round 2.  The cases still run in ONE child process (one torch import, one CUDA context; a faulting
kernel would poison the CUDA context of the process that launched it and must not take the rest of the suite down)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CASES = ['reference_fixture', 'oracle_default_shape', 'oracle_ragged_many_slots', 'oracle_no_confounders',
         'oracle_ipsmf_exposure', 'equals_general_scorer_eval_batch', 'out_of_range_ids',
         'projected_noise_equals_reference_formula', 'projected_noise_same_distribution',
         'device_confounder_draw_equals_torch_randint', 'predict_many_with_device_confounders']


@pytest.fixture(scope='module')
def worker_output():
    """All cases in ONE child process (one torch import, one CUDA context); each prints `GATHER_OK <case>` or
    `GATHER_FAIL <case>: <why>`.  A hard fault ends the child: the cases after it then report 'not reached'."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__)] + CASES, capture_output=True, text=True,
                           timeout=900, cwd=ROOT)
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or '')
        return -9, out, 'child timed out after 900 s'
    return r.returncode, r.stdout, r.stderr


@pytest.mark.parametrize('case', CASES)
def test_gather_scorer(worker_output, case):
    rc, out, err = worker_output
    failed = [ln for ln in out.splitlines() if ln.startswith('GATHER_FAIL ' + case)]
    assert not failed, failed[0]
    assert 'GATHER_OK ' + case in out, 'not reached (child rc=%d): %s' % (rc, (err or out)[-2000:])


# ---------------------------------------------------------------------------------------------------------
# worker side (python tests/test_gpu_zz_gather.py <case>)
# ---------------------------------------------------------------------------------------------------------
def _run(case):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import golden_params, rel_err, GOLDEN
    from oracle import dccf_oracle as O
    from test_gpu_parity import make_model, random_problem

    def fd_of(X, si):
        return {'X': torch.from_numpy(X).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0,
                'sample_item': torch.from_numpy(si)}

    def gather_predict(model, X, si):
        from dccf_b200 import kernels
        model.use_gather_scorer = True
        calls, orig = [], kernels.score_gather
        kernels.score_gather = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
        try:
            out = model.predict(fd_of(X, si))['prediction']
        finally:
            kernels.score_gather = orig
        torch.cuda.synchronize()
        model.check_ids()
        assert len(calls) == (1 if len(X) else 0)          # the gather kernel scored the batch, not the general scorer
        return out.cpu().numpy()

    if case == 'reference_fixture':
        # the reference's own predictions with --std 0 --dropout 0 and A = 3 (tests/golden/train_nodrop.npz)
        g = np.load(os.path.join(GOLDEN, 'train_nodrop.npz'), allow_pickle=False)
        assert float(g['std']) == 0.0
        model = make_model(golden_params(g, 'final_'), int(g['S']), int(g['A']), 0.0)
        got = gather_predict(model, g['eval_X'], g['eval_sample_item'])               # ragged: 37 pairs
        assert rel_err(got, g['eval_pred']) < 1e-5
        model = make_model(golden_params(g), int(g['S']), int(g['A']), 0.0)
        got = gather_predict(model, g['X_0'], g['sample_item_0'])
        assert rel_err(got, g['pred_0']) < 1e-5
    elif case in ('oracle_default_shape', 'oracle_ragged_many_slots', 'oracle_no_confounders'):
        U, I, F, P, S, A = {'oracle_default_shape': (300, 500, 768, 1001, 10, 2),
                            'oracle_ragged_many_slots': (50, 70, 128, 37, 40, 1),      # Z = 41: three slot chunks
                            'oracle_no_confounders': (20, 30, 64, 3, 0, 2)}[case]
        params, X, si, _, _ = random_problem(21, U, I, F, P, S, A, 0.0, 0.0)
        model = make_model(params, S, A, 0.0)
        ref = O.predict(params, X, si, None, None, A, dtype=np.float64)
        got = gather_predict(model, X, si)
        assert got.shape == (P,) and rel_err(got, ref['pred']) < 1e-5
        one = gather_predict(model, X[:1], si[:1])                                      # a single pair
        assert rel_err(one, ref['pred'][:1]) < 1e-5
    elif case == 'oracle_ipsmf_exposure':
        from dccf_b200 import synth
        U, I, F, P, S, A = 80, 120, 64, 64, 10, 2
        params, X, si, _, _ = random_problem(7, U, I, F, P, S, A, 0.0, 0.0)
        fac = synth.make_ipsmf_factors(U, I, seed=3)
        model = make_model(params, S, A, 0.0, expo_factors=fac)
        ref = O.predict(params, X, si, None, None, A, dtype=np.float64, expo=fac)
        assert rel_err(gather_predict(model, X, si), ref['pred']) < 1e-5
    elif case == 'equals_general_scorer_eval_batch':
        # a full evaluation batch (16 384 pairs, 1 001 candidates per user) against the general FP32 kernel
        U, I, F, S, A = 500, 3000, 768, 10, 2
        params, _, _, _, _ = random_problem(3, U, I, F, 2, S, A, 0.0, 0.0)
        rs = np.random.RandomState(8)
        P = 16384
        X = np.stack([np.repeat(np.arange(U), 1001)[:P], rs.randint(0, I, P)], 1).astype(np.int64)
        si = rs.randint(0, I, (P, S)).astype(np.int64)
        model = make_model(params, S, A, 0.0)
        model.use_gather_scorer = False
        want = model.predict(fd_of(X, si))['prediction'].cpu().numpy()
        got = gather_predict(model, X, si)
        assert rel_err(got, want) < 1e-5
        # after a parameter update the tables are rebuilt
        with torch.no_grad():
            model.mlp[0].weight.mul_(1.5)
            model.iid_embeddings.weight.add_(0.01)
        model.use_gather_scorer = False
        want2 = model.predict(fd_of(X, si))['prediction'].cpu().numpy()
        got2 = gather_predict(model, X, si)
        assert rel_err(got2, want2) < 1e-5 and rel_err(got2, want) > 1e-3
    elif case == 'out_of_range_ids':
        params, X, si, _, _ = random_problem(4, 40, 60, 64, 10, 3, 1, 0.0, 0.0)
        model = make_model(params, 3, 1, 0.0)
        model.use_gather_scorer = True
        bad = si.copy()
        bad[2, 1] = 60
        model.predict(fd_of(X, bad))
        try:
            model.check_ids()
        except IndexError:
            pass
        else:
            raise AssertionError('an item id outside the table must raise')
        assert gather_predict(model, X[:0], si[:0]).shape == (0,)                       # empty batch
    elif case == 'projected_noise_equals_reference_formula':
        # eval_noise = 'projected' draws g ~ N(0, std^2 I_64) per row (the library's Philox stream, materialised here by
        # dccf_noise_fill for the same seed / offset) and adds M·g to the pre-activation, M·M^T = W_f·W_f^T.  The same
        # prediction must come out of the REFERENCE formula (oracle, float64) fed the 768-d noise tensor
        # eps = W_f^+ · M · g  (W_f^+ the pseudo-inverse, so that W_f·eps = M·g).
        from dccf_b200 import kernels
        U, I, F, P, S, A, std = 200, 300, 768, 300, 10, 2, 0.1
        params, X, si, _, _ = random_problem(31, U, I, F, P, S, A, 0.0, 0.0)
        model = make_model(params, S, A, std, seed=77)
        model.eval_noise = 'projected'
        N = P * (S + 1) * A
        model._rng_offset = 10                                  # the call below uses offset 11
        got = model.predict(dict(fd_of(X, si), force_tc=True))['prediction'].cpu().numpy()
        model.check_ids()
        g = torch.empty((N, 64), device='cuda')
        kernels.noise_fill(g, std, 77, 11)
        g = g.cpu().numpy().astype(np.float64)
        Wf = params['W'][:, 64:].astype(np.float64)
        M = model.projection_factor().numpy()
        assert np.abs(M @ M.T - Wf @ Wf.T).max() < 1e-12 * np.abs(Wf @ Wf.T).max() + 1e-18
        pinv = Wf.T @ np.linalg.inv(Wf @ Wf.T)                  # [768, 64]
        eps = g @ M.T @ pinv.T                                  # [N, 768]:  W_f · eps_r = M · g_r
        ref = O.predict(params, X, si, eps, None, A, dtype=np.float64)
        assert rel_err(got, ref['pred']) < 1e-5
        # explicit noise tensors keep the exact formulation whatever eval_noise says
        noise = (np.random.RandomState(1).standard_normal((N, F)) * std).astype(np.float32)
        fd = dict(fd_of(X, si), force_tc=True, noise=torch.from_numpy(noise))
        ref = O.predict(params, X, si, noise, None, A, dtype=np.float64)
        assert rel_err(model.predict(fd)['prediction'].cpu().numpy(), ref['pred']) < 1e-5
    elif case == 'projected_noise_same_distribution':
        # exact and projected draws are different random numbers with the same law: per-pair mean and variance of the
        # prediction over 64 independent calls agree within sampling error (averaged over 2 048 pairs)
        U, I, F, P, S, A, std = 300, 500, 768, 2048, 10, 2, 0.3
        params, X, si, _, _ = random_problem(9, U, I, F, P, S, A, 0.0, 0.0)
        params['W'] = (params['W'] * 4).astype(np.float32)       # make the noise term matter
        model = make_model(params, S, A, std)
        stats = {}
        for mode in ('exact', 'projected'):
            model.eval_noise = mode
            draws = torch.stack([model.predict(dict(fd_of(X, si), force_tc=True))['prediction'] for _ in range(64)])
            stats[mode] = (draws.mean(0).double().cpu().numpy(), draws.var(0).double().cpu().numpy())
        mean_e, var_e = stats['exact']
        mean_p, var_p = stats['projected']
        assert var_e.mean() > 0 and var_p.mean() > 0
        assert abs(var_p.mean() / var_e.mean() - 1.0) < 0.05                       # 64 x 2048 samples: ~1 % noise
        # per-pair means differ by sampling noise only: |diff| ~ sqrt(2 var / 64)
        zscore = (mean_e - mean_p) / np.sqrt((var_e + var_p) / 64 + 1e-30)
        assert abs(zscore.mean()) < 0.2 and 0.8 < zscore.std() < 1.2
    elif case == 'device_confounder_draw_equals_torch_randint':
        # k_confounder_draw continues torch's CPU generator on the device: the ids torch.randint would have produced,
        # bit for bit, across calls of awkward sizes, and torch continues correctly from the state written back
        from dccf_b200 import host_rng
        for seed, high in ((2019, 16000), (3, 1), (11, 999983), (5, (1 << 28) - 1)):
            a, b = torch.Generator(), torch.Generator()
            a.manual_seed(seed)
            b.manual_seed(seed)
            if seed == 11:                                   # start in the middle of a generation
                torch.randint(10, (100,), generator=a)
                torch.randint(10, (100,), generator=b)
            stream = host_rng.DeviceStream(torch.device('cuda'), generator=b)
            for shape in ((7,), (16384, 10), (0, 10), (3, 1), (624,), (1, 623), (37, 10)):
                want = torch.randint(high, shape, generator=a)
                got = stream.draw(high, shape)
                assert got.dtype == torch.int64 and tuple(got.shape) == tuple(shape)
                assert torch.equal(got.cpu(), want), (seed, high, shape)
            stream.finish()
            assert torch.equal(torch.randint(1 << 20, (1000,), generator=a), torch.randint(1 << 20, (1000,), generator=b))
            assert torch.equal(torch.randn(4, generator=a), torch.randn(4, generator=b))
    elif case == 'predict_many_with_device_confounders':
        # an evaluation pass with the confounders drawn on the device == the same pass with host draws: identical
        # predictions (same ids, same noise offsets) and an identical torch generator afterwards
        U, I, F, S, A, std = 200, 300, 128, 10, 2, 0.1
        params, _, _, _, _ = random_problem(13, U, I, F, 2, S, A, 0.0, 0.0)
        rs = np.random.RandomState(2)
        sizes = [4096, 4096, 333]
        Xs = [np.stack([rs.randint(0, U, n), rs.randint(0, I, n)], 1).astype(np.int64) for n in sizes]
        results = []
        for device_draw in (False, True):
            model = make_model(params, S, A, std)
            model.device_confounders = device_draw
            torch.manual_seed(99)
            fds = [{'X': torch.from_numpy(x).cuda(), 'rank': 1, 'train': False, 'dropout': 0.0} for x in Xs]
            outs = [o.cpu().numpy() for o in model.predict_many(fds)]
            model.check_ids()
            results.append((outs, torch.get_rng_state().clone(), torch.randint(1 << 20, (50,))))
        for a, b in zip(results[0][0], results[1][0]):
            assert np.array_equal(a, b)
        assert torch.equal(results[0][2], results[1][2])
    else:
        raise SystemExit('unknown case ' + case)
    print('GATHER_OK ' + case)


if __name__ == '__main__':
    import traceback
    sys.path.insert(0, ROOT)
    for name in sys.argv[1:]:
        try:
            _run(name)
        except Exception as exc:        # noqa: BLE001 — reported per case; a sticky CUDA error fails the later ones too
            traceback.print_exc()
            print('GATHER_FAIL %s: %s' % (name, str(exc).splitlines()[0][:300] if str(exc) else type(exc).__name__))
