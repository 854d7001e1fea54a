"""Host side of the multi-GPU path on CPU: world_size-2 gloo exchange of the packed gradient segments,
user sharding of the evaluation set, metric all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dccf_b200.dist import (GradExchange, IdExchange, all_reduce_sum, rank_slice_of_draws, shard_users,
                            step_partition, sync_host_rng)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        P, Z, D, K = 6, 3, 64, 128
        ex = GradExchange(P, Z, D, K, world, rank, torch.device('cpu'))
        v = ex.send_views()
        g = torch.Generator().manual_seed(100 + rank)
        v['gu_rec'].copy_(torch.randn(P, D, generator=g))
        v['gi_rec'].copy_(torch.randn(P * Z, D, generator=g))
        v['gW'].copy_(torch.randn(D, K, generator=g))
        v['gb'].copy_(torch.randn(D, generator=g))
        v['keys_u'].copy_(torch.arange(P, dtype=torch.int32) + 1000 * rank)
        v['keys_i'].copy_(torch.arange(P * Z, dtype=torch.int32) + 5000 * rank)
        v['loss'].fill_(float(rank + 1))
        ex.exchange_records()              # the two segments travel separately (records first, dW / db / loss later)
        ex.exchange_dense()
        # every rank sees every segment, rank-major, keys intact as int32
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            assert torch.equal(ex.recv_part('gu', r).view(P, D), torch.randn(P, D, generator=gr))
            assert torch.equal(ex.recv_part('gi', r).view(P * Z, D), torch.randn(P * Z, D, generator=gr))
            assert torch.equal(ex.recv_part('gW', r).view(D, K), torch.randn(D, K, generator=gr))
            assert torch.equal(ex.recv_part('keys_u', r), torch.arange(P, dtype=torch.int32) + 1000 * r)
            assert torch.equal(ex.recv_part('keys_i', r), torch.arange(P * Z, dtype=torch.int32) + 5000 * r)
        assert float(ex.total_loss()) == sum(range(1, world + 1))
        assert ex.rec.seg % 4 == 0 and ex.dense.seg % 4 == 0 and all(a % 4 == 0 for a, _ in ex.off.values())
        ex.done()
        # the ids of every rank's batch, gathered at the start of a step: int64 views of a float32 segment
        S = 5
        ix = IdExchange(P, S, world, rank, torch.device('cpu'))
        ix.send_X.copy_(torch.arange(2 * P, dtype=torch.int64).view(P, 2) + (1 << 40) * (rank + 1))
        ix.send_si.copy_(torch.arange(P * S, dtype=torch.int64).view(P, S) + 7 * rank)
        ids = ix.exchange().view(torch.int64)
        for r in range(world):
            seg = ids[r * ix.seg_i64:(r + 1) * ix.seg_i64]
            assert torch.equal(seg[:2 * P].view(P, 2), torch.arange(2 * P, dtype=torch.int64).view(P, 2) + (1 << 40) * (r + 1))
            assert torch.equal(seg[ix.si_off_i64:ix.si_off_i64 + P * S].view(P, S),
                               torch.arange(P * S, dtype=torch.int64).view(P, S) + 7 * r)
        ix.done()
        # evaluation: users partitioned, metric sums reduced
        rs = np.random.RandomState(0)
        uid = rs.randint(0, 37, size=500)
        mine = shard_users(uid, rank, world)
        sums = all_reduce_sum([float(len(mine)), float(len(set(uid[mine].tolist())))])
        assert sums[0] == 500 and sums[1] == len(set(uid.tolist()))
        np.save(os.path.join(out_dir, 'rows_%d.npy' % rank), mine)
        # host generators drift apart during sharded evaluation and are put back in step before training
        class _M(object):
            _rng_offset = 0
        m = _M()
        torch.manual_seed(5)
        torch.randint(10, (100 * (rank + 1),))          # ranks consume different amounts
        m._rng_offset = 40 + rank
        sync_host_rng(m)
        torch.save({'draw': torch.randint(1 << 20, (8,)), 'offset': m._rng_offset},
                   os.path.join(out_dir, 'rng_%d.pt' % rank))
        # row-sharded user table: the checkpoint holds the FULL table under the reference's keys (rank 0 collects the
        # rows), and every rank loads its own slice back
        from dccf_b200.models.DCCF import DCCF
        U, I, F = 10, 7, 64
        lo, hi = (0, 6) if rank == 0 else (6, 10)
        path = os.path.join(out_dir, 'sharded.pt')
        model = DCCF(path='', dataset='', sentence_model='', sample_num=2, attribute_num=1, std=0.0, label_min=0,
                     label_max=1, feature_num=0, user_num=U, item_num=I, u_vector_size=64, i_vector_size=64, n_layers=1,
                     random_seed=1, model_path=path, feature_embedding=torch.zeros(I, F), expo_prob=torch.ones(U, I),
                     user_shard=(lo, hi))
        with torch.no_grad():
            model.uid_embeddings.weight.copy_(torch.arange(lo, hi, dtype=torch.float32)[:, None].expand(hi - lo, 64))
            model.iid_embeddings.weight.fill_(3.0)
        assert model.expo_prob.shape[0] == hi - lo
        model.save_model()
        sd = torch.load(path, map_location='cpu')
        assert sd['uid_embeddings.weight'].shape == (U, 64)
        assert torch.equal(sd['uid_embeddings.weight'][:, 0], torch.arange(U, dtype=torch.float32))
        assert sorted(sd) == ['iid_embeddings.weight', 'mlp.0.bias', 'mlp.0.weight', 'uid_embeddings.weight']
        with torch.no_grad():
            model.uid_embeddings.weight.zero_()
        model.load_model()
        assert torch.equal(model.uid_embeddings.weight[:, 5], torch.arange(lo, hi, dtype=torch.float32))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_and_sharding(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rs = np.random.RandomState(0)
    uid = rs.randint(0, 37, size=500)
    rows = [np.load(os.path.join(str(tmp_path), 'rows_%d.npy' % r)) for r in range(2)]
    assert len(np.intersect1d(rows[0], rows[1])) == 0
    assert len(rows[0]) + len(rows[1]) == 500
    # a user's candidates never straddle two ranks
    assert len(set(uid[rows[0]].tolist()) & set(uid[rows[1]].tolist())) == 0


    a, b = (torch.load(os.path.join(str(tmp_path), 'rng_%d.pt' % r)) for r in range(2))
    assert torch.equal(a['draw'], b['draw']) and a['offset'] == b['offset'] == 40


def test_step_partition_covers_every_batch_once():
    """src/main.py under torchrun: the global step k = batches k*world .. k*world+world-1, the leftover full batches
    form the replicated tail; ranks' confounder slices re-assemble the single generator stream."""
    for n_full, world in ((6750, 8), (13, 4), (3, 4), (8, 2), (0, 2)):
        seen = []
        tails = set()
        for r in range(world):
            mine, tail = step_partition(n_full, world, r)
            assert len(mine) == n_full // world
            seen += mine
            tails.add(tuple(tail))
        assert len(tails) == 1                                       # every rank agrees on the tail
        tail = list(tails.pop())
        assert sorted(seen + tail) == list(range(n_full)) and len(tail) < world
    g = torch.Generator().manual_seed(3)
    m, world, P, S = 5, 4, 6, 10
    draws = torch.randint(1000, (m * world, P, S), generator=g)
    parts = [rank_slice_of_draws(draws, world, r) for r in range(world)]
    for k in range(m):
        for r in range(world):
            assert torch.equal(parts[r][k], draws[k * world + r])


def test_shard_users_single_rank_and_balance():
    uid = np.repeat(np.arange(100), 11)
    assert np.array_equal(shard_users(uid, 0, 1), np.arange(1100))
    sizes = [len(shard_users(uid, r, 8)) for r in range(8)]
    assert sum(sizes) == 1100 and max(sizes) - min(sizes) <= 11
