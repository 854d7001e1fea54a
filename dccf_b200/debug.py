"""Debug instrumentation: per-kernel start / end times inside a CUDA-graph replay of the fused training step.

The kernels of the step record {earliest CTA start, latest CTA end} in nanoseconds of %globaltimer into a device
buffer installed with dccf_debug_timeline_* (include/dccf_b200.h).  CUDA events cannot bracket the nodes of a graph;
these stamps are taken by the kernels' own CTAs on the stream(s) the graph runs them on."""
import ctypes

import numpy as np
import torch

from . import _lib

SLOT_NAMES = ['k_link_ids', 'k_adam_untouched', 'k_train_fwd_tc', 'k_train_mid', 'k_train_bwd_tc', 'k_adam_touched',
              'k_stage_batch', 'fwd latest CTA start', 'k_dp_push ids', 'k_dp_wait', 'k_dp_push records', 'k_dp_push dense',
              'peers arrived (k_adam_touched)', 'k_csr_build', 'dense push: pieces issued', 'dense push: copies complete']
N_SLOTS = 16


class StepTimeline(object):
    def __init__(self, device):
        self.lib = _lib.load()
        self.slots = torch.zeros(2 * N_SLOTS, dtype=torch.int64, device=device)
        self.init = torch.tensor([-1, 0] * N_SLOTS, dtype=torch.int64, device=device)     # -1 = UINT64_MAX for atomicMin
        self.rows = []

    def __enter__(self):
        for fn in (self.lib.dccf_debug_timeline_train, self.lib.dccf_debug_timeline_adam, self.lib.dccf_debug_timeline_dp):
            _lib.check(fn(ctypes.c_void_p(self.slots.data_ptr())), 'dccf_debug_timeline')
        return self

    def __exit__(self, *a):
        torch.cuda.synchronize()
        for fn in (self.lib.dccf_debug_timeline_train, self.lib.dccf_debug_timeline_adam, self.lib.dccf_debug_timeline_dp):
            fn(None)

    def arm(self):
        """Reset the slots (stream-ordered, no host synchronisation): the next step is the one recorded."""
        self.slots.copy_(self.init, non_blocking=True)

    def collect(self):
        torch.cuda.synchronize()
        self.rows.append(self.slots.cpu().numpy().astype(np.uint64).reshape(N_SLOTS, 2).copy())

    def summary(self):
        """{kernel: {'start_us', 'end_us', 'us'}} (medians over the collected steps, relative to the first kernel
        start of the step) and the median step length."""
        rows = np.stack(self.rows)
        used = [i for i in range(len(SLOT_NAMES)) if i != 7 and rows[0, i, 1] != 0 and rows[0, i, 0] != np.uint64(2**64 - 1)]
        if not used:
            return {}, 0.0
        t0 = np.array([min(int(r[i, 0]) for i in used) for r in rows], dtype=np.float64)
        out = {}
        for i in used:
            st = (rows[:, i, 0].astype(np.float64) - t0) / 1e3
            en = (rows[:, i, 1].astype(np.float64) - t0) / 1e3
            out[SLOT_NAMES[i]] = {'start_us': float(np.median(st)), 'end_us': float(np.median(en)),
                                  'us': float(np.median(en - st))}
        step_us = float(np.median([(max(int(r[i, 1]) for i in used) - t) / 1e3 for r, t in zip(rows, t0)]))
        return out, step_us
