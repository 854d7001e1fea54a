"""Timeline of the kernels inside one CUDA-graph replay of the fused training step (dccf_b200/debug.py).  Prints,
per kernel, the median start / end / duration in microseconds relative to the first kernel start of the step, with
and without an L2 flush before the step.  Usage (B200): python tools/step_timeline.py [--steps 40]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dccf_b200.debug import StepTimeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=40)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    U, I = bench.PRESETS['electronics']
    model = bench.build_model(U, I, dev)
    n = 3 * args.steps + 10
    X = bench.synth_batches(n, U, I, 3)
    X_dev = torch.from_numpy(X).to(dev)
    si_dev = torch.randint(I, size=(n, 2 * bench.BATCH, bench.S)).to(dev)
    step = model.begin_resident_epoch(X_dev, si_dev, bench.DROPOUT)
    for _ in range(8):
        step()
    flush = bench.L2Flusher(dev)
    for mode in ('back to back (L2 warm)', 'L2 flushed before the step'):
        with StepTimeline(dev) as tl:
            for _ in range(args.steps // 2):
                # two untimed steps first, no host synchronisation in between: the measured step starts on a busy GPU
                step()
                step()
                if mode.startswith('L2'):
                    flush()
                tl.arm()
                step()
                tl.collect()
            kernels, step_us = tl.summary()
        print('--- %s: median over %d steps, microseconds from the first kernel start ---' % (mode, len(tl.rows)))
        for name, k in sorted(kernels.items(), key=lambda kv: kv[1]['start_us']):
            print('%-18s start %7.2f  end %7.2f  dur %6.2f' % (name, k['start_us'], k['end_us'], k['us']))
        print('step end %.2f' % step_us)


if __name__ == '__main__':
    main()
