"""BaseRunner.fit's batch schedule on CPU, with a recording stand-in for the model: which batches and which confounder
draws reach the device-resident epoch / the step-by-step tail, on one process and as a rank of a data-parallel job
(src/main.py under torchrun: --batch_size per rank, global step k = batches k*world .. k*world + world - 1)."""
import numpy as np
import pytest
import torch

from dccf_b200.runners.BaseRunner import BaseRunner


class _OnDevice(torch.Tensor):
    """A CPU tensor that claims to live on the GPU (fit only takes the resident-epoch path for CUDA batches)."""
    is_cuda = property(lambda self: True)


def _dev(t):
    return torch.Tensor._make_subclass(_OnDevice, t)


class _Recorder(object):
    sample_num = 10

    def __init__(self, item_num=1000, dp=None):
        self.item_num, self._dp, self.optimizer = item_num, dp, object()
        self.resident, self.stepwise, self.suspended_flags = [], [], []

    def train(self):
        pass

    def eval(self):
        pass

    def draw_confounders(self, n_pairs):
        return torch.randint(self.item_num, (n_pairs, self.sample_num))

    def begin_resident_epoch(self, X_epoch, draws, dropout):
        state = {'k': 0}
        n = X_epoch.shape[0]

        def step():
            k = state['k']
            state['k'] += 1
            self.resident.append((torch.Tensor(X_epoch[k]).clone(), draws[k].clone()))
            return {'prediction': torch.zeros(1), 'check': [], 'loss': torch.zeros(())}
        step.remaining = lambda: n - state['k']
        step.first = None
        return step

    def train_step(self, fd):
        self.stepwise.append((torch.Tensor(fd['X']).clone(), fd['sample_item'].clone()))
        self.suspended_flags.append(self._dp is None)
        return {'prediction': torch.zeros(1), 'check': [], 'loss': torch.zeros(())}

    def data_parallel_suspended(self):
        import contextlib

        @contextlib.contextmanager
        def cm():
            saved, self._dp = self._dp, None
            try:
                yield self
            finally:
                self._dp = saved
        return cm()


class _DP(object):
    rank = 1

    def __init__(self, batches):
        self._b = batches

    def prepare_batches(self, data, batch_size, train):
        return [dict(b) for b in self._b]


def _batches(n_full, ragged, P=8):
    out = []
    for k in range(n_full):
        out.append({'X': _dev(torch.full((P, 2), k, dtype=torch.int64)), 'Y': torch.zeros(P)})
    if ragged:
        out.append({'X': _dev(torch.full((ragged, 2), n_full, dtype=torch.int64)), 'Y': torch.zeros(ragged)})
    return out


def _run(n_full, ragged, world, rank, monkeypatch):
    runner = BaseRunner(optimizer='Adam', learning_rate=1e-3, metrics='ndcg@5', dropout=0.2)
    runner.show_progress = False
    model = _Recorder(dp={'rank': rank, 'world': world} if world > 1 else None)
    monkeypatch.setattr(BaseRunner, '_world', staticmethod(lambda m: world))
    torch.manual_seed(7)
    runner.fit(model, data=None, data_processor=_DP(_batches(n_full, ragged)))
    return model


@pytest.mark.parametrize('n_full,ragged', [(5, 3), (4, 0), (1, 0), (0, 5), (2, 1)])
def test_single_process_schedule_is_the_reference_order(n_full, ragged, monkeypatch):
    """Every batch once, in order; the confounder draws are the single torch.randint stream of DCCF.py:72."""
    m = _run(n_full, ragged, 1, 0, monkeypatch)
    seen = m.resident + m.stepwise
    assert [int(x[0, 0]) for x, _ in seen] == list(range(n_full + (1 if ragged else 0)))
    torch.manual_seed(7)
    sizes = [8] * n_full + ([ragged] if ragged else [])
    for (x, si), P in zip(seen, sizes):
        assert x.shape[0] == P
    stream = torch.randint(1000, (sum(sizes), 10))
    assert torch.equal(torch.cat([si for _, si in seen]), stream)


@pytest.mark.parametrize('n_full,ragged,world', [(9, 3, 2), (8, 0, 4), (3, 2, 4), (13, 0, 8)])
def test_data_parallel_schedule(n_full, ragged, world, monkeypatch):
    """As rank r of `world`: global step k takes batch k*world + r from the resident epoch; the batches that do not
    fill a global step run step by step with data parallelism suspended, on every rank alike; all ranks consume the
    generator identically and the ranks' draws re-assemble the single stream."""
    runs = [_run(n_full, ragged, world, r, monkeypatch) for r in range(world)]
    n_steps = n_full // world
    torch.manual_seed(7)
    stream = torch.randint(1000, (n_steps * world * 8, 10)).view(n_steps * world, 8, 10)
    for r, m in enumerate(runs):
        assert [int(x[0, 0]) for x, _ in m.resident] == [k * world + r for k in range(n_steps)]
        for k, (_, si) in enumerate(m.resident):
            assert torch.equal(si, stream[k * world + r])
        tail = list(range(n_steps * world, n_full)) + ([n_full] if ragged else [])
        assert [int(x[0, 0]) for x, _ in m.stepwise] == tail
        assert all(m.suspended_flags) and m._dp is not None            # suspended inside, restored after
    for a, b in zip(runs[0].stepwise, runs[-1].stepwise):              # replicated tail: identical inputs on every rank
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
