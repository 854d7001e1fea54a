"""Legs of bench.py beyond the headline train / eval pair: one function per object on the JSON line.

    dp_parity_check        N > 1, before any timing: the data-parallel step (eager AND CUDA-graph replay, the library's
                           own noise / dropout streams) equals the single-GPU step on the concatenated batch; replicas
                           bit-identical.  bench.py exits non-zero when it fails.
    full_catalogue_leg     BASELINE.json configs[3]: Yelp shape, full-catalogue scoring on the tcgen05 GEMM + fused top-k
                           (deterministic DCCF predictor and the IPSBiasedMF exposure matrix), users sharded over N ranks
    scaled_leg             BASELINE.json configs[4]: 10 M users x 1 M items, user table / Adam state / IPS-MF user factors
                           row-sharded over N ranks, exposure on the fly; equality check at a small shape + ms / step
    gpu_eager_reference    SURVEY.md §8d "the thing to beat": the reference's op sequence (oracle/torch_port.py) run
                           eagerly by PyTorch on the same B200 — train step and one 16 384-pair evaluation batch
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = 0x9E3779B97F4A7C15


def _dist():
    import torch.distributed as dist
    return dist


def _max_over_ranks(x, world, dev):
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    _dist().all_reduce(t, op=_dist().ReduceOp.MAX)
    return float(t.item())


def _all_ok(ok, world, dev):
    if world == 1:
        return bool(ok)
    t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
    _dist().all_reduce(t, op=_dist().ReduceOp.MIN)
    return bool(t.item() > 0.5)


def _small_model(params, S, A, std, dev, expo_factors=None, user_shard=None, user_num=None):
    from dccf_b200.models.DCCF import DCCF
    U = user_num if user_num is not None else params['E_user'].shape[0]
    I = params['E_item'].shape[0]
    model = DCCF(path='', dataset='', sentence_model='', sample_num=S, attribute_num=A, std=std, label_min=0, label_max=1,
                 feature_num=0, user_num=U, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1, random_seed=2019,
                 model_path='/tmp/dccf_bench_check.pt', feature_embedding=params['Feat'],
                 expo_prob=None if expo_factors is not None else params['expo'], expo_factors=expo_factors,
                 user_shard=user_shard)
    with torch.no_grad():
        eu = params['E_user'] if user_shard is None else params['E_user'][user_shard[0]:user_shard[1]]
        model.uid_embeddings.weight.copy_(eu)
        model.iid_embeddings.weight.copy_(params['E_item'])
        model.mlp[0].weight.copy_(params['W'])
        model.mlp[0].bias.copy_(params['b'])
    return model.to(dev)


def _params_of(model):
    return {'E_user': model.uid_embeddings.weight.detach(), 'E_item': model.iid_embeddings.weight.detach(),
            'W': model.mlp[0].weight.detach(), 'b': model.mlp[0].bias.detach()}


def _err98(a, b):
    """(98th percentile, max) of |a - b| / max|b|.  Data-parallel ranks and the single GPU tile the two contractions
    differently, so a pre-activation within 1e-7 of zero can pass the ReLU on one side only and move the entries it
    feeds: the bulk must agree to fp32 accuracy, single entries may differ (a wrong segment, a missing rank or a
    misplaced record moves most entries by far more)."""
    e = (a.double() - b.double()).abs() / b.double().abs().max().clamp_min(1e-30)
    return float(torch.quantile(e.flatten()[:16_000_000], 0.98)), float(e.max())


def dp_parity_check(world, rank, dev, steps=4, user_sharded=False):
    """The data-parallel training step against ONE GPU on the concatenated batch, on the library's own random streams
    (each rank's Philox streams are materialised with dccf_noise_fill / dccf_dropout_mask_fill and handed to the
    single-GPU model as explicit tensors): step 1 runs kernel by kernel, steps 2.. replay the captured CUDA graph with
    the peer-memory exchange inside.  user_sharded: the scaled configuration (row-sharded user table, on-the-fly
    IPSBiasedMF exposure, user-row gradients never leave their rank)."""
    from dccf_b200 import kernels, synth
    dist = _dist()
    S, A, std, drop, D = 10, 2, 0.1, 0.2, 64
    U, I, F = (64 * world, 300, 128) if user_sharded else (300, 400, 768)
    b_loc = 32
    P_loc, P, R = 2 * b_loc, 2 * b_loc * world, (S + 1) * A
    b = P // 2
    g = torch.Generator().manual_seed(17 + (100 if user_sharded else 0))         # identical on every rank
    params = {'E_user': torch.randn((U, D), generator=g) * 0.05, 'E_item': torch.randn((I, D), generator=g) * 0.05,
              'W': torch.randn((D, D + F), generator=g) * 0.05, 'b': torch.randn(D, generator=g) * 0.05,
              'Feat': torch.randn((I, F), generator=g) / F ** 0.5, 'expo': torch.rand((U, I), generator=g)}
    per = U // world
    if user_sharded:      # rank r's positives (and their negatives) carry rank r's users
        u = torch.cat([torch.randint(r * per, (r + 1) * per, (b_loc,), generator=g) for r in range(world)])
    else:
        u = torch.randint(0, U, (b,), generator=g)
    X = torch.stack([torch.cat([u, u]), torch.randint(0, I, (P,), generator=g)], 1).contiguous()
    si = torch.randint(0, I, (steps, P, S), generator=g)
    pairs_of = [torch.cat([torch.arange(r * b_loc, (r + 1) * b_loc), b + torch.arange(r * b_loc, (r + 1) * b_loc)])
                for r in range(world)]
    rows_of = [(p[:, None] * R + torch.arange(R)[None, :]).reshape(-1).to(dev) for p in pairs_of]
    fac = None
    if user_sharded:
        fac = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else float(v))
               for k, v in synth.make_ipsmf_factors(U, I, seed=3).items()}
    shard = (rank * per, (rank + 1) * per) if user_sharded else None
    model = _small_model(params, S, A, std, dev, expo_factors=fac, user_shard=shard, user_num=U)
    model.enable_data_parallel()
    model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
    single = _small_model(params, S, A, std, dev, expo_factors=fac, user_num=U)
    single.optimizer = single.make_fused_optimizer(lr=1e-3, l2=1e-4)
    mine = pairs_of[rank]
    Y = torch.zeros(P, device=dev)
    info = {'world': world, 'steps': steps, 'pairs_per_rank': P_loc, 'user_sharded': bool(user_sharded)}
    ok, why = True, ''
    try:
        loss_err = 0.0
        for t in range(steps):
            off = model._rng_offset + 1                      # the per-call counter this step's streams will use
            out = model.train_step({'X': X[mine].to(dev), 'rank': 1, 'train': True, 'dropout': drop, 'Y': Y[:P_loc],
                                    'sample_item': si[t][mine]})
            noise = torch.empty((P * R, F), device=dev)
            mask = torch.empty((P * R, D), device=dev)
            for r in range(world):
                seed_r = (2019 + GOLD * r) & 0xffffffffffffffff
                n_r = torch.empty((P_loc * R, F), device=dev)
                m_r = torch.empty((P_loc * R, D), device=dev)
                kernels.noise_fill(n_r, std, seed_r, off)
                kernels.dropout_mask_fill(m_r, drop, seed_r, off)
                noise[rows_of[r]] = n_r
                mask[rows_of[r]] = m_r
            out1 = single.train_step({'X': X.to(dev), 'rank': 1, 'train': True, 'dropout': drop, 'Y': Y,
                                      'sample_item': si[t], 'noise': noise, 'dropout_mask': mask})
            l, l1 = float(out['loss']), float(out1['loss'])
            loss_err = max(loss_err, abs(l - l1) / max(abs(l1), 1e-30))
        model.check_ids()
        info['exchange'] = model._exchange_for(P_loc).mode
        info['graph_steps'] = steps - 1 if model.__dict__.get('_graphs') else 0
        info['loss_rel_err'] = loss_err
        got, want = _params_of(model), _params_of(single)
        errs = {}
        for k in ('E_user', 'E_item', 'W', 'b'):
            w = want[k][shard[0]:shard[1]] if (user_sharded and k == 'E_user') else want[k]
            errs[k] = _err98(got[k], w)
        info['err_p98_max'] = {k: [float('%.3g' % v[0]), float('%.3g' % v[1])] for k, v in errs.items()}
        tol = {'E_user': 1e-5, 'E_item': 1e-5, 'b': 1e-5, 'W': 5e-4}
        for k, (q, mx) in errs.items():
            if not (q < tol[k] and mx < 2e-2):
                ok, why = False, '%s: 98th percentile %.3g, max %.3g' % (k, q, mx)
        if not loss_err < 1e-5:
            ok, why = False, 'loss differs by %.3g' % loss_err
        # replicated tensors bit-identical on every rank
        same = True
        for k in (('E_item', 'W', 'b') if user_sharded else ('E_user', 'E_item', 'W', 'b')):
            ref = got[k].clone()
            dist.broadcast(ref, 0)
            same = same and bool(torch.equal(ref, got[k]))
        info['replicas_identical'] = bool(_all_ok(same, world, dev))
        if not info['replicas_identical']:
            ok, why = False, 'replicas diverged'
    except Exception as exc:        # noqa: BLE001 — reported; bench.py exits non-zero
        ok, why = False, '%s: %s' % (type(exc).__name__, str(exc)[:300])
    info['ok'] = bool(_all_ok(ok, world, dev))
    if why:
        info['why'] = why
    return info


def replica_checksum(model, world, dev, skip_user_table=False):
    """After the timed steps: every rank's parameters hash to the same value (int32 wrap-around sum of the raw bits)."""
    if world == 1:
        return True
    sums = []
    for k, v in _params_of(model).items():
        if skip_user_table and k == 'E_user':
            continue
        sums.append(v.contiguous().view(torch.int32).sum(dtype=torch.int64))
    t = torch.stack(sums).double()
    lo, hi = t.clone(), t.clone()
    _dist().all_reduce(lo, op=_dist().ReduceOp.MIN)
    _dist().all_reduce(hi, op=_dist().ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


# ------------------------------------------------------------------------------------------------------------------
# configs[3]: full-catalogue scoring, Yelp shape, user-sharded
# ------------------------------------------------------------------------------------------------------------------
def full_catalogue_leg(world, rank, dev, peaks, U=32000, I=38000, k=5, iters=10, warmup=3, F=768):
    """Each rank owns a contiguous block of users (no data-path collective); CUDA events per iteration, L2 flushed
    between iterations, max over ranks.  (a) deterministic DCCF predictor over the whole catalogue with fused top-k:
    score[u,i] = <E_user[u], relu(W_i E_item[i] + W_f Feat[i] + b)> — item-side projection (dccf_tc_prepare) + GEMM;
    (b) the IPSBiasedMF exposure matrix (src/models/IPSBiasedMF.py:37-57), top-k only and materialised.  Algorithmic
    flops = 2 x U x I x 64 (the tensor cores issue 3 TF32 products per FP32 product: frac_issued)."""
    from dccf_b200 import full_catalogue, synth
    hbm_peak, bf16_peak, peak_src = peaks
    tf32_peak = bf16_peak / 2.0
    lo, hi = U * rank // world, U * (rank + 1) // world
    Ul = hi - lo
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def timed(fn):
        ms = []
        for i in range(warmup + iters):
            flush_buf.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                ms.append(e0.elapsed_time(e1))
        return _max_over_ranks(float(np.mean(ms)), world, dev)

    out = {'metric': 'full_catalogue_users_per_s', 'unit': 'users/s', 'n_gpus': world, 'scaling': 'strong',
           'config': {'workload': 'full-catalogue scoring + fused top-%d on synthetic yelp-shape data: U=%d I=%d D=64 F=%d, '
                                  'users sharded over %d rank(s), no data-path collective' % (k, U, I, F, world),
                      'l2_cache': 'flushed between timed iterations (256 MiB write)'}}
    # (b) IPSBiasedMF exposure
    fac = synth.make_ipsmf_factors(U, I, seed=2019)
    f = {kk: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else float(v)) for kk, v in fac.items()}
    f['mf_user'] = f['mf_user'][lo:hi].contiguous()
    f['mf_user_bias'] = f['mf_user_bias'][lo:hi].contiguous()
    topk_ms = timed(lambda: full_catalogue.ipsmf_topk(f, k))
    mat_ms = timed(lambda: full_catalogue.ipsmf_exposure(f))
    # parity of the fused top-k against the materialised matrix (first 512 users of the shard)
    mat = full_catalogue.ipsmf_exposure(f)
    ts, ti = full_catalogue.ipsmf_topk(f, k)
    n_chk = min(512, Ul)
    want_s, want_i = torch.topk(mat[:n_chk], k, dim=1)
    parity = bool(torch.equal(ts[:n_chk], want_s)) and float((ti[:n_chk] != want_i).float().mean()) < 1e-3
    del mat
    flop = 2.0 * Ul * I * 64
    out['ipsmf_exposure'] = {
        'topk_ms': topk_ms, 'materialise_ms': mat_ms, 'topk_matches_materialised': _all_ok(parity, world, dev),
        'users_per_s_topk': U / (topk_ms / 1e3),
        'roofline_topk': {'kernel': 'k_full_scores (+ k_topk_merge)', 'bound': 'tensor',
                          'achieved': flop / 1e12 / (topk_ms / 1e3), 'peak': tf32_peak, 'unit': 'TFLOP/s',
                          'frac': flop / 1e12 / (topk_ms / 1e3) / tf32_peak,
                          'frac_issued': 3 * flop / 1e12 / (topk_ms / 1e3) / tf32_peak,
                          # ncu --set full, one launch on one GPU (profiles/r2_ncu_full_catalogue_summary.csv): the user
                          # factors and the pre-split item images in, k ids out — no re-reads from DRAM
                          'traffic': 30.27e6 if (world == 1 and U == 32000 and I == 38000) else None,
                          'peak_source': peak_src + ': bf16 burst / 2 for kind::tf32'},
        'roofline_materialise': {'kernel': 'k_full_scores', 'bound': 'hbm', 'achieved': Ul * I * 4 / 1e9 / (mat_ms / 1e3),
                                 'peak': hbm_peak, 'unit': 'GB/s', 'frac': Ul * I * 4 / 1e9 / (mat_ms / 1e3) / hbm_peak,
                                 'note': 'algorithmic bytes = 4 B per score written (per rank)'}}
    del f
    # (a) deterministic DCCF predictor over the catalogue
    g = torch.Generator(device='cpu').manual_seed(2024)
    params = {'E_user': torch.randn((Ul, 64), generator=torch.Generator().manual_seed(7 + rank)) * 0.05,
              'E_item': torch.randn((I, 64), generator=g) * 0.05, 'W': torch.randn((64, 64 + F), generator=g) * 0.05,
              'b': torch.randn(64, generator=g) * 0.05, 'Feat': torch.randn((I, F), generator=g) / F ** 0.5,
              'expo': torch.ones((1, 1))}
    from dccf_b200.models.DCCF import DCCF
    model = DCCF(path='', dataset='', sentence_model='', sample_num=0, attribute_num=1, std=0.0, label_min=0, label_max=1,
                 feature_num=0, user_num=Ul, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1, random_seed=2019,
                 model_path='/tmp/dccf_bench_fc.pt', feature_embedding=params['Feat'], expo_prob=torch.ones((Ul, 1)))
    with torch.no_grad():
        model.uid_embeddings.weight.copy_(params['E_user'])
        model.iid_embeddings.weight.copy_(params['E_item'])
        model.mlp[0].weight.copy_(params['W'])
        model.mlp[0].bias.copy_(params['b'])
    model = model.to(dev)

    def dccf_pass():
        model._param_epoch += 1            # the item-side projection is part of every pass (parameters change per epoch)
        return full_catalogue.dccf_catalogue_topk(model, k)

    dccf_ms = timed(dccf_pass)
    warm_ms = timed(lambda: full_catalogue.dccf_catalogue_topk(model, k))     # projected tables cached
    flop_proj = 2.0 * I * (64 + F) * 64
    out['dccf_catalogue'] = {
        'pass_ms': dccf_ms, 'gemm_topk_ms': warm_ms, 'users_per_s': U / (dccf_ms / 1e3),
        'note': 'pass = item-side projection of the catalogue (dccf_tc_prepare: I x (64 + F) x 64) + relu + U x I x 64 '
                'GEMM with fused per-user top-%d; gemm_topk_ms = the GEMM + top-k alone (tables cached)' % k,
        'roofline': {'kernel': 'k_full_scores (+ k_topk_merge)', 'bound': 'tensor',
                     'achieved': flop / 1e12 / (warm_ms / 1e3), 'peak': tf32_peak, 'unit': 'TFLOP/s',
                     'frac': flop / 1e12 / (warm_ms / 1e3) / tf32_peak,
                     'frac_issued': 3 * flop / 1e12 / (warm_ms / 1e3) / tf32_peak,
                     'pass_frac': (flop + flop_proj) / 1e12 / (dccf_ms / 1e3) / tf32_peak,
                     'peak_source': peak_src + ': bf16 burst / 2 for kind::tf32'}}
    out['value'] = U / (dccf_ms / 1e3)
    del model
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# configs[4]: scaled shape, row-sharded user table, on-the-fly exposure
# ------------------------------------------------------------------------------------------------------------------
def scaled_leg(world, rank, dev, peaks, U=10_000_000, I=1_000_000, F=768, steps=20, warmup=5):
    """Row-sharded 10 M x 1 M configuration: the user table, its Adam state and the IPS-MF user factors
    hold only this rank's users; the item table, W, the features and the IPS-MF item factors are replicated; exposure
    is evaluated on the fly from the IPSBiasedMF formula (no user x item matrix: 40 TB at this shape); each rank trains
    on its own users, so user-row gradients never leave the rank.  Reports the equality check at a small shape
    (N > 1: against one unsharded GPU) and ms / step at the full shape; the step is bounded by the dense l2 + clip +
    Adam sweep over every row (24 B per parameter per step, src/runners/BaseRunner.py:100,181,187)."""
    from dccf_b200.models.DCCF import DCCF
    hbm_peak, _, peak_src = peaks
    out = {'metric': 'train_samples_per_s', 'unit': 'samples/s', 'n_gpus': world, 'scaling': 'weak',
           'config': {'workload': 'scaled synthetic DCCF: U=%d I=%d D=64 F=%d, user table row-sharded over %d rank(s) '
                                  '(%d users each), exposure on the fly from IPSBiasedMF factors, batch 128 per rank'
                                  % (U, I, F, world, U // world)}}
    if world > 1:
        out['parity'] = dp_parity_check(world, rank, dev, steps=3, user_sharded=True)
    per = U // world
    lo, hi = rank * per, (rank + 1) * per
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    gi = torch.Generator(device=dev).manual_seed(7)                 # replicated tensors: same seed on every rank

    def rnd(shape, gen, scale):
        return torch.randn(shape, generator=gen, device=dev) * scale

    fac = {'mf_user': rnd((per, 64), g, 0.1), 'mf_item': rnd((I, 64), gi, 0.1), 'mf_user_bias': rnd((per,), g, 0.1),
           'mf_item_bias': rnd((I,), gi, 0.1), 'mf_global_bias': 0.1,
           'propensity': torch.rand((I,), generator=gi, device=dev), 'mf_min_propensity': 0.1}
    feat = rnd((I, F), gi, F ** -0.5)
    torch.manual_seed(2019)
    t_build = time.perf_counter()
    # (the module is built on the device: init_paras' N(0, 0.01) of a 10^7-row table on the CPU generator takes tens
    # of seconds and is not what this leg measures)
    with torch.device(dev):
        model = DCCF(path='', dataset='', sentence_model='', sample_num=10, attribute_num=2, std=0.1, label_min=0,
                     label_max=1, feature_num=0, user_num=U, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1,
                     random_seed=2019, model_path='/tmp/dccf_scaled.pt', feature_embedding=feat, expo_prob=None,
                     expo_factors=fac, user_shard=(lo, hi))
        gw = torch.Generator(device=dev).manual_seed(11)            # item table / W / b identical on every rank
        with torch.no_grad():
            model.uid_embeddings.weight.copy_(rnd((per, 64), g, 0.01))
            model.iid_embeddings.weight.copy_(rnd((I, 64), gw, 0.01))
            model.mlp[0].weight.copy_(rnd((64, 64 + F), gw, 0.01))
            model.mlp[0].bias.copy_(rnd((64,), gw, 0.01))
    model = model.to(dev)
    if world > 1:
        model.enable_data_parallel()
    model.optimizer = model.make_fused_optimizer(lr=1e-3, l2=1e-4)
    n = steps + warmup
    gh = torch.Generator().manual_seed(rank)
    u = torch.randint(lo, hi, (n, 128), generator=gh)
    it = torch.randint(0, I, (n, 256), generator=gh)
    X_epoch = torch.stack([torch.cat([u, u], 1), it], 2).to(dev)                  # [n, 256, 2]
    si_epoch = torch.randint(0, I, (n, 256, 10), generator=gh).to(dev)
    step = model.begin_resident_epoch(X_epoch, si_epoch, 0.2)
    if step is None:
        raise RuntimeError('the CUDA-graph step is unavailable (peer-memory exchange?)')
    while step.remaining() > steps:
        step()
    torch.cuda.synchronize()
    if world > 1:
        _dist().barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = 0
    while step.remaining() > 0:
        o = step()
        k += 1
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1) / max(k, 1), world, dev)
    model.check_ids()
    loss = float(o['loss'])
    sweep_gb = 24.0 * ((per + I) * 64 + 64 * (64 + F) + 64) / 1e9
    out.update({'value': world * 128 / (ms / 1e3), 'ms_per_step': ms, 'steps': k, 'last_loss': loss,
                'loss_finite': bool(np.isfinite(loss)),
                'replicas_identical': replica_checksum(model, world, dev, skip_user_table=True),
                'hbm_allocated_gb': torch.cuda.max_memory_allocated() / 1e9, 'build_s': time.perf_counter() - t_build,
                'roofline': {'kernel': 'k_adam_untouched (wide) + k_adam_touched', 'bound': 'hbm',
                             'achieved': sweep_gb / (ms / 1e3), 'peak': hbm_peak, 'unit': 'GB/s',
                             'frac': sweep_gb / (ms / 1e3) / hbm_peak, 'traffic': None, 'peak_source': peak_src,
                             'note': 'algorithmic bytes = 24 B per parameter of this rank per step (%.2f GB) over the WHOLE '
                                     'step time: the dense l2 + clip + Adam sweep the reference semantics require bounds '
                                     'the step at this shape; timed back to back, tables (>= 2 GB) larger than L2' % sweep_gb}})
    del model, step, X_epoch, si_epoch, fac, feat
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------------------
# the reference's op sequence, eager PyTorch on the same GPU
# ------------------------------------------------------------------------------------------------------------------
def gpu_eager_reference(U, I, dev, cfg, steps=30, warmup=5):
    """oracle/torch_port.py (the reference's ATen calls in the reference's order, src/models/DCCF.py:66-127 and
    src/runners/BaseRunner.py:175-188) on the GPU: ms per 128-sample training step and per 16 384-pair evaluation
    batch, wall clock with a synchronise on both sides (the eager path is bound by its ~100 launches per step and by
    the temporaries it materialises).  The confounder draw stays on the CPU generator and is copied, as in the
    reference.  Timed, not shipped: nothing on the product path imports oracle/."""
    from oracle import torch_port
    D, F, S, A = cfg['D'], cfg['F'], cfg['S'], cfg['A']
    gen = torch.Generator(device='cpu').manual_seed(cfg['seed'] + 5)
    feat = torch.randn((I, F), generator=gen) / (F ** 0.5)
    expo = torch.rand((U, I), generator=torch.Generator(device=dev).manual_seed(cfg['seed'] + 6), device=dev)
    model = torch_port.DCCFPort(U, I, feat, torch.zeros((1, 1)), sample_num=S, attribute_num=A, std=cfg['std'],
                                seed=cfg['seed'])
    model.to_device(dev)
    model.expo_prob = expo
    optim = torch.optim.Adam(model.parameters(), lr=cfg['lr'], weight_decay=cfg['l2'])
    B = cfg['batch']
    Y = torch.cat([torch.ones(B), torch.zeros(B)]).to(dev)
    rs = np.random.RandomState(cfg['seed'])
    model.train()
    ts = []
    for i in range(warmup + steps):
        u = rs.randint(0, U, B)
        X = torch.from_numpy(np.stack([np.concatenate([u, u]), rs.randint(0, I, 2 * B)], 1)).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch_port.fit_step(model, optim, {'X': X, 'Y': Y, 'dropout': cfg['dropout']}, cfg['l2'])
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    train_ms = 1e3 * float(np.mean(ts))
    model.eval()
    Pe = cfg['eval_batch']
    Xe = torch.stack([torch.randint(0, U, (Pe,)), torch.randint(0, I, (Pe,))], 1).to(dev)
    te = []
    with torch.no_grad():
        for i in range(2 + 3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            model.predict({'X': Xe, 'dropout': 0.0})
            torch.cuda.synchronize()
            if i >= 2:
                te.append(time.perf_counter() - t0)
    eval_ms = 1e3 * float(np.mean(te))
    peak_gb = torch.cuda.max_memory_allocated() / 1e9
    del model, optim, expo
    torch.cuda.empty_cache()
    return {'kind': 'eager PyTorch port of the reference ops on the same GPU (oracle/torch_port.py, fp32, TF32 off)',
            'train_ms_per_step': train_ms, 'train_samples_per_s': B / (train_ms / 1e3),
            'eval_ms_per_batch': eval_ms, 'eval_pairs_per_batch': Pe,
            'eval_users_per_s': (Pe / float(1 + cfg['test_neg_n'])) / (eval_ms / 1e3),
            'eval_note': 'scoring only: the reference then ranks on the host with pandas (src/models/BaseModel.py:83-112)',
            'hbm_allocated_gb': peak_gb, 'steps': steps, 'warmup': warmup}
