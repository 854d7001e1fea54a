"""Device time of the evaluation ranker (dccf_rank_eval_multi) alone: n_users x 1001 candidates, metrics at k = 5.

    python tools/rank_bench.py [--users 1024 16384] [--k 5]

CUDA events around each launch, L2 flushed (256 MiB write) before it, median of 20; prints one JSON line per size with
the time, GB/s on 8 B per candidate (score + label, the compulsory bytes) and on 20 B (ids and the row indirection
counted as the round-1 kernel read them), for the grouped layout (no cand_rows) and the indirect one."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dccf_b200 import kernels  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, nargs='+', default=[1024, 16384])
    ap.add_argument('--k', type=int, default=5)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    for n_users in args.users:
        n_c = 1001
        rows = n_users * n_c
        g = torch.Generator(device=dev).manual_seed(3)
        scores = torch.randn(rows, generator=g, device=dev)
        labels = torch.zeros(rows, device=dev)
        labels[::n_c] = 1.0
        iids = torch.randint(0, 16000, (rows,), generator=g, device=dev)
        off = torch.arange(0, rows + 1, n_c, dtype=torch.int64, device=dev)
        ident = torch.arange(rows, dtype=torch.int32, device=dev)
        sums = torch.empty((1, 5), dtype=torch.float64, device=dev)
        out = {}
        for name, cand in (('grouped', None), ('indirect', ident)):
            # the launch is replayed from a CUDA graph: the kernel is shorter than the Python path to it
            kernels.rank_eval_multi(scores, labels, iids, cand, off, [args.k], out_sums=sums)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                kernels.rank_eval_multi(scores, labels, iids, cand, off, [args.k], out_sums=sums)
            ts = []
            for it in range(25):
                flush_buf.fill_(float(it))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                if it >= 5:
                    ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            out[name] = {'ms': ms, 'gbs_8B': rows * 8 / 1e9 / (ms / 1e3), 'gbs_20B': rows * 20 / 1e9 / (ms / 1e3)}
        print(json.dumps({'users': n_users, 'candidates': n_c, 'k': args.k, **out,
                          'ndcg_sum': float(sums[0, 0].item())}))


if __name__ == '__main__':
    main()
