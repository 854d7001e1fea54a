"""Data-parallel plumbing (new work — the reference is single-process).

One process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  Training is
data parallel with replicated parameters: the BPR loss is a SUM over samples (src/models/DCCF.py:120), so
gradients add across ranks.  Per step each rank contributes its ids (start of the step), its gradient records
(after the middle kernel) and its dW / db / loss (after the dW kernel) as packed segments; an all-gather delivers
every rank's segments to every rank; each rank then applies the identical l2 + clip + Adam update over the
identical record list (segment-major, fixed order), so replicas stay bit-identical without any parameter broadcast.  Evaluation shards users across ranks and needs one small
all-reduce of the metric sums.
"""
import numpy as np
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class SegmentExchange(object):
    """All-gather of one packed float32 segment per rank: `send` [seg] -> `recv` [world, seg] on every rank.
    Peer-memory mode (default on CUDA): the rank's own kernels store the segment into every peer's symmetric
    buffer over NVLink and wait on arrival flags (csrc/dp_exchange.cu) — plain kernels, capturable in a CUDA
    graph, on whatever stream is current.  Otherwise one all_gather_into_tensor."""

    def __init__(self, seg, world, rank, device, group=None, use_p2p=True):
        self.seg = (int(seg) + 3) // 4 * 4
        self.world, self.rank, self.group = world, rank, group
        self.send = torch.zeros(self.seg, dtype=torch.float32, device=device)
        self.mode = 'collective'
        self.recv = None
        # the peer-memory protocol keeps int32[8] arrival / consumed flags (DP_MAX_WORLD in csrc/dp_exchange.cu): larger
        # worlds (more than one NVSwitch domain of 8) use the collective
        if world > 1 and world <= 8 and torch.device(device).type == 'cuda' and use_p2p:
            try:
                self._init_p2p(device)
                self.mode = 'p2p'
            except Exception as e:                      # symmetric memory unavailable: NCCL all-gather instead
                import logging
                logging.warning('dccf_b200.dist: peer-memory exchange unavailable (%s); using all_gather' % (e,))
        if self.recv is None:
            self.recv = torch.zeros(world * self.seg, dtype=torch.float32, device=device)

    def _init_p2p(self, device):
        """Symmetric receive buffer mapped on every peer (torch symmetric memory = CUDA VMM/IPC over NVLink):
        [ recv: world x seg floats | flag area: an arrival flag per (source rank, CTA of its push), consumed flags ]."""
        import torch.distributed._symmetric_memory as symm
        from . import kernels               # (fails early if the library is missing)
        self.flag_off = self.world * self.seg
        buf = symm.empty(self.flag_off + (kernels.dp_flag_floats() + 3) // 4 * 4, dtype=torch.float32, device=device)
        hdl = symm.rendezvous(buf, dist.group.WORLD if self.group is None else self.group)
        buf.zero_()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)                  # every rank's flags are zero before anyone pushes
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        delta = int(buf.data_ptr()) - ptrs[self.rank]   # offset of `buf` inside the symmetric allocation
        self.peer_bases = [p + delta for p in ptrs]
        self.sym, self.hdl = buf, hdl
        self.recv = buf[:self.flag_off]
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.cta_counter = torch.zeros(1, dtype=torch.int32, device=device)

    def push(self, folds=None):
        """Peer-memory mode: store this rank's segment into every peer's receive buffer and publish its arrival — no
        wait: the consumer kernel waits for the peers' segments itself (kernels.make_dp_sync) or `wait` is called."""
        from . import kernels
        if folds:
            kernels.dp_push_fold(self.send, self.seg, self.peer_bases, self.world, self.rank, self.flag_off,
                                 self.epoch_dev, self.cta_counter, folds)
        else:
            kernels.dp_push(self.send, self.seg, self.peer_bases, self.world, self.rank, self.flag_off,
                            self.epoch_dev, self.cta_counter)

    def wait(self):
        from . import kernels
        kernels.dp_wait(self.sym, self.seg, self.world, self.flag_off, self.epoch_dev)

    def exchange(self, folds=None):
        """All ranks' segments, rank-major, in self.recv.  folds (peer-memory mode only): ranges of the segment that
        are summed from partial buffers on the way out instead of being read from `send`:
        [(parts, n_parts, stride, offset_in_segment, n_floats), ...] (csrc/dp_exchange.cu)."""
        if self.mode == 'p2p':
            self.push(folds)
            self.wait()
        elif self.world == 1:
            self.recv.copy_(self.send)
        else:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        return self.recv

    def done(self):
        """The receive buffer has been consumed (call after the last kernel that reads it)."""
        if self.mode == 'p2p':
            from . import kernels
            kernels.dp_done(self.peer_bases, self.world, self.rank, self.flag_off, self.epoch_dev)


class IdExchange(SegmentExchange):
    """The ids of every rank's batch, gathered at the START of a step (24 KB per rank): with them each rank knows
    which table rows the global step touches before any gradient exists, so the Adam sweep of all other rows can
    overlap the forward and backward (DCCF._fused_split_step).  Segment = [X int64 [P,2] | sample_item int64 [P,S]];
    `send_X` / `send_si` are the step's input buffers themselves (no staging copy)."""

    def __init__(self, P, S, world, rank, device, group=None, use_p2p=True):
        self.P, self.S = P, S
        SegmentExchange.__init__(self, 2 * (2 * P) + 2 * (P * S), world, rank, device, group, use_p2p)
        self.send_X = self.send[:4 * P].view(torch.int64).view(P, 2)
        self.send_si = self.send[4 * P:4 * P + 2 * P * S].view(torch.int64).view(P, S)
        self.si_off_i64 = 2 * P                          # int64 elements from a segment's X to its sample_item
        self.seg_i64 = self.seg // 2

    def recv_i64(self):
        return self.recv.view(torch.int64)


class GradExchange(object):
    """The per-rank gradient contribution of a step, all-gathered in TWO packed segments:

        records [ user-row gradient records | item-row gradient records | user keys | item keys ]   (ready after the
                 middle kernel of the step; shipped on a side stream while the dW kernel runs)
        dense   [ dW | db | loss ]                                                                  (after the dW kernel)

    P pairs per rank and step, Z slots, D embedding width, K = D + F columns of W.  Offsets are in 4-byte elements
    inside their segment; the key arrays are int32 views of the same float32 buffer."""

    def __init__(self, P, Z, D, K, world, rank, device, group=None, use_p2p=True, user_records=True):
        self.P, self.Z, self.D, self.K = P, Z, D, K
        self.world, self.rank = world, rank
        # user_records=False: the user table is row-sharded and each rank only trains its own users, so the
        # user-row gradients never leave the rank (scaled config)
        self.user_records = user_records
        Pu = P if user_records else 0
        self.off, self.where = {}, {}
        sizes = {}
        for which, parts in (('rec', (('gu', Pu * D), ('gi', P * Z * D), ('keys_u', Pu), ('keys_i', P * Z))),
                             ('dense', (('gW', D * K), ('gb', D), ('loss', 1)))):
            off = 0
            for name, n in parts:
                self.off[name] = (off, n)
                self.where[name] = which
                off += (n + 3) // 4 * 4                  # keep every part 16-byte aligned
            sizes[which] = off
        self.rec = SegmentExchange(sizes['rec'], world, rank, device, group, use_p2p)
        self.dense = SegmentExchange(sizes['dense'], world, rank, device, group, use_p2p)
        self.mode = 'p2p' if (self.rec.mode == 'p2p' and self.dense.mode == 'p2p') else 'collective'

    def _ex(self, name):
        return self.rec if self.where[name] == 'rec' else self.dense

    def seg_of(self, name):
        """segment length (floats) of the exchange that carries `name`"""
        return self._ex(name).seg

    def send_part(self, name):
        a, n = self.off[name]
        t = self._ex(name).send[a:a + n]
        return t.view(torch.int32) if name.startswith('keys') else t

    def recv_part(self, name, seg_index=0):
        """part `name` of rank seg_index's segment in the receive buffer; the parts of consecutive ranks are
        seg_of(name) elements apart"""
        a, n = self.off[name]
        ex = self._ex(name)
        t = ex.recv[seg_index * ex.seg + a: seg_index * ex.seg + a + n]
        return t.view(torch.int32) if name.startswith('keys') else t

    def send_views(self):
        D = self.D
        return {'gu_rec': self.send_part('gu').view(-1, D), 'gi_rec': self.send_part('gi').view(self.P * self.Z, D),
                'gW': self.send_part('gW').view(D, self.K), 'gb': self.send_part('gb'),
                'keys_u': self.send_part('keys_u'), 'keys_i': self.send_part('keys_i'), 'loss': self.send_part('loss')}

    def exchange_records(self):
        return self.rec.exchange()

    def exchange_dense(self, folds=None):
        if folds and self.dense.mode == 'p2p':
            return self.dense.exchange(folds=folds)
        return self.dense.exchange()

    def done(self):
        self.rec.done()
        self.dense.done()

    def total_loss(self):
        a, _ = self.off['loss']
        return self.dense.recv.view(self.world, self.dense.seg)[:, a].sum()


def step_partition(n_full, world, rank):
    """Which of an epoch's `n_full` equal-size batches a rank trains on (src/main.py under torchrun; --batch_size is
    then PER RANK: the global step k is made of the batches k*world .. k*world + world - 1, one per rank — weak scaling,
    SURVEY.md §8e (ii)).  Returns (mine, tail): `mine[k]` = index of this rank's batch in global step k; `tail` = the
    full batches that do not fill a last global step — every rank runs those (and the ragged last batch) itself,
    replicated, so no sample of the epoch is dropped."""
    n_steps = n_full // world
    return [k * world + rank for k in range(n_steps)], list(range(n_steps * world, n_full))


def rank_slice_of_draws(draws, world, rank):
    """draws [m * world, P, S] = the confounder draws of m global steps in batch order (ONE generator stream, consumed
    identically on every rank so the ranks' generators stay in step) -> [m, P, S], this rank's batches."""
    return draws.view(-1, world, draws.shape[1], draws.shape[2])[:, rank]


def sync_host_rng(model, src=0):
    """Put every rank's torch CPU generator and the model's noise / dropout call counter where rank `src` has them.
    Sharded evaluation consumes both by different amounts on different ranks; training needs them in step (every rank
    draws the global step's confounders and keeps its slice; replicated steps must be bit-identical)."""
    if not is_distributed():
        return
    box = [torch.get_rng_state(), int(model._rng_offset)] if dist.get_rank() == src else [None, None]
    dist.broadcast_object_list(box, src=src)
    torch.set_rng_state(box[0])
    model._rng_offset = int(box[1])


def init_from_env():
    """(rank, local_rank, world) from torchrun's environment; initialises NCCL and selects the GPU when world > 1."""
    import os
    rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not (dist.is_available() and dist.is_initialized()):
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, local, world


def shard_users(uid, rank, world):
    """Row indices of the contiguous block of (sorted) users owned by `rank`, balanced by row count; all
    candidates of a user stay on one rank."""
    uid = np.asarray(uid)
    order = np.argsort(uid, kind='stable')
    users, starts = np.unique(uid[order], return_index=True)
    bounds = np.concatenate([starts, [len(uid)]])
    target = [len(uid) * r // world for r in range(world + 1)]
    cut = np.searchsorted(bounds, target, side='left')
    cut[0], cut[-1] = 0, len(users)
    lo, hi = bounds[cut[rank]], bounds[cut[rank + 1]]
    return np.sort(order[lo:hi])


def all_reduce_sum(values):
    """Sum a small float64 vector over ranks (metric sums, user counts)."""
    t = torch.as_tensor(values, dtype=torch.float64)
    if is_distributed():
        if dist.get_backend() == 'nccl':
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
