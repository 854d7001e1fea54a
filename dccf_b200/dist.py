"""Data-parallel plumbing (new work — the reference is single-process).

One process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  Training is
data parallel with replicated parameters: the BPR loss is a SUM over samples (src/models/DCCF.py:120), so
gradients add across ranks.  Per step each rank contributes ONE packed segment

    [ user-row gradient records | item-row gradient records | dW | db | user keys | item keys | loss ]

and a single all-gather delivers every rank's segment to every rank; each rank then applies the identical
dense l2 + clip + Adam sweep over the identical record list (segment-major, fixed order), so replicas stay
bit-identical without any parameter broadcast.  Evaluation shards users across ranks and needs one small
all-reduce of the metric sums.
"""
import numpy as np
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GradExchange(object):
    """Layout of the per-rank gradient segment and the all-gather of all segments.

    P pairs per rank and step, Z slots, D embedding width, K = D + F columns of W.  Offsets are in 4-byte
    elements; the key arrays are int32 views of the same float32 buffer."""

    def __init__(self, P, Z, D, K, world, rank, device, group=None):
        self.P, self.Z, self.D, self.K = P, Z, D, K
        self.world, self.rank, self.group = world, rank, group
        off = 0
        self.off = {}
        for name, n in (('gu', P * D), ('gi', P * Z * D), ('gW', D * K), ('gb', D), ('keys_u', P),
                        ('keys_i', P * Z), ('loss', 1)):
            self.off[name] = (off, n)
            off += (n + 3) // 4 * 4                      # keep every part 16-byte aligned
        self.seg = off
        self.recv = torch.zeros(world * self.seg, dtype=torch.float32, device=device)
        # the send segment is this rank's slot of the receive buffer when the backend allows in-place
        # all-gather; kept separate for portability (gloo)
        self.send = torch.zeros(self.seg, dtype=torch.float32, device=device)

    def part(self, buf, name, seg_index=0):
        a, n = self.off[name]
        t = buf[seg_index * self.seg + a: seg_index * self.seg + a + n]
        if name.startswith('keys'):
            t = t.view(torch.int32)
        return t

    def send_views(self):
        D = self.D
        return {'gu_rec': self.part(self.send, 'gu').view(self.P, D),
                'gi_rec': self.part(self.send, 'gi').view(self.P * self.Z, D),
                'gW': self.part(self.send, 'gW').view(D, self.K), 'gb': self.part(self.send, 'gb'),
                'keys_u': self.part(self.send, 'keys_u'), 'keys_i': self.part(self.send, 'keys_i'),
                'loss': self.part(self.send, 'loss')}

    def exchange(self):
        """All ranks' segments, rank-major, in self.recv."""
        if self.world == 1:
            self.recv.copy_(self.send)
        else:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        return self.recv

    def total_loss(self):
        a, _ = self.off['loss']
        return self.recv.view(self.world, self.seg)[:, a].sum()


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
