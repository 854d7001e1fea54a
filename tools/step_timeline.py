"""Timeline of the kernels inside one CUDA-graph replay of the fused training step (debug instrumentation:
dccf_debug_timeline_*, %globaltimer per CTA).  Prints, per kernel, the median start / end / duration in
microseconds relative to the first kernel start of the step, with and without an L2 flush before the step.
Usage (B200): python tools/step_timeline.py [--steps 40]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dccf_b200 import _lib  # noqa: E402

NAMES = ['k_link_ids', 'k_adam_untouched', 'k_train_fwd_tc', 'k_train_mid', 'k_train_bwd_tc', 'k_adam_touched',
         'k_stage_batch', 'fwd latest CTA start']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=40)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    U, I = bench.PRESETS['electronics'][:2] if hasattr(bench, 'PRESETS') else (48000, 16000)
    model = bench.build_model(U, I, dev)
    n = 3 * args.steps + 10
    rs = np.random.RandomState(3)
    b = bench.BATCH
    u = rs.randint(0, U, size=(n, b))
    X = np.concatenate([np.stack([u, rs.randint(0, I, size=(n, b))], 2), np.stack([u, rs.randint(0, I, size=(n, b))], 2)],
                       axis=1).astype(np.int64)
    X_dev = torch.from_numpy(X).to(dev)
    si_dev = torch.randint(I, size=(n, 2 * b, bench.S)).to(dev)
    step = model.begin_resident_epoch(X_dev, si_dev, bench.DROPOUT)
    for _ in range(8):
        step()
    torch.cuda.synchronize()
    lib = _lib.load()
    slots = torch.zeros(16, dtype=torch.int64, device=dev)
    for fn in (lib.dccf_debug_timeline_train, lib.dccf_debug_timeline_adam):
        _lib.check(fn(ctypes.c_void_p(slots.data_ptr())), 'dccf_debug_timeline')
    init = torch.tensor([-1, 0] * 8, dtype=torch.int64, device=dev)      # -1 = UINT64_MAX for the atomicMin
    flush = bench.L2Flusher(dev)
    for mode in ('back to back (L2 warm)', 'L2 flushed before the step'):
        rows = []
        for _ in range(args.steps // 2):
            # two untimed steps first, no host synchronisation in between: the measured step starts on a busy GPU
            # (after an idle gap the first kernel of a replay runs ~9 us slower — clocks / caches of an idle chip)
            step()
            step()
            if mode.startswith('L2'):
                flush()
            slots.copy_(init, non_blocking=True)
            step()
            torch.cuda.synchronize()
            rows.append(slots.cpu().numpy().astype(np.uint64).reshape(8, 2))
        rows = np.stack(rows)                                              # [steps, 8, 2]
        used = [i for i in range(7) if rows[0, i, 1] != 0]
        t0 = np.array([min(int(r[i, 0]) for i in used) for r in rows], dtype=np.float64)
        print('--- %s: median over %d steps, microseconds from the first kernel start ---' % (mode, len(rows)))
        order = sorted(used, key=lambda i: np.median(rows[:, i, 0].astype(np.float64) - t0))
        for i in order:
            st = (rows[:, i, 0].astype(np.float64) - t0) / 1e3
            en = (rows[:, i, 1].astype(np.float64) - t0) / 1e3
            print('%-18s start %7.2f  end %7.2f  dur %6.2f' % (NAMES[i], np.median(st), np.median(en), np.median(en - st)))
        print('fwd: latest CTA start %.2f' % np.median((rows[:, 7, 1].astype(np.float64) - t0) / 1e3))
        print('step end %.2f' % np.median([(max(int(r[i, 1]) for i in used) - t) / 1e3 for r, t in zip(rows, t0)]))
    for fn in (lib.dccf_debug_timeline_train, lib.dccf_debug_timeline_adam):
        fn(None)


if __name__ == '__main__':
    main()
