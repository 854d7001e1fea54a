"""Full-catalogue scoring on the tcgen05 GEMM (BASELINE.json config 4: Yelp-shape, user-sharded across ranks).

    python tools/bench_full_catalogue.py [--users 32000 --items 38000 --k 5 --iters 20]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_full_catalogue.py

Each rank owns a contiguous block of users (no data-path collective).  Two measurements, CUDA events, L2 flushed
between iterations: (1) fused top-k only — the [U,I] matrix is never written; (2) materialised matrix — the
IPSBiasedMF exposure matrix the reference stores as <ds>.ips_expo_prob.npy.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--users', type=int, default=32000)
    ap.add_argument('--items', type=int, default=38000)
    ap.add_argument('--k', type=int, default=5)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    from dccf_b200 import full_catalogue, synth
    lo, hi = a.users * rank // world, a.users * (rank + 1) // world
    fac = synth.make_ipsmf_factors(a.users, a.items, seed=2019)
    f = {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else float(v)) for k, v in fac.items()}
    f['mf_user'] = f['mf_user'][lo:hi].contiguous()
    f['mf_user_bias'] = f['mf_user_bias'][lo:hi].contiguous()
    U = hi - lo
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def timed(fn):
        ms = []
        for i in range(a.warmup + a.iters):
            flush_buf.fill_(float(i))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if i >= a.warmup:
                ms.append(e0.elapsed_time(e1))
        t = float(np.mean(ms))
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=dev)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            t = float(tt.item())
        return t

    topk_ms = timed(lambda: full_catalogue.ipsmf_topk(f, a.k))
    mat_ms = timed(lambda: full_catalogue.ipsmf_exposure(f)) if U * a.items * 4 < 40e9 else None
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
    tf32_peak = peaks['bf16_tflops'] / 2
    flop3 = 3 * 2.0 * U * a.items * 64           # the three TF32 products issued
    line = {'metric': 'full_catalogue_users_per_s', 'value': a.users / (topk_ms / 1e3), 'unit': 'users/s',
            'n_gpus': world, 'config': {'workload': 'IPSBiasedMF full-catalogue scoring + fused top-%d, U=%d I=%d D=64, '
                                        'user-sharded' % (a.k, a.users, a.items)},
            'topk_ms': topk_ms, 'dtype': 'f32 (3xTF32 on tcgen05)',
            'roofline_topk': {'bound': 'tensor', 'achieved': flop3 / 1e12 / (topk_ms / 1e3), 'peak': tf32_peak,
                              'unit': 'TFLOP/s', 'frac': flop3 / 1e12 / (topk_ms / 1e3) / tf32_peak},
            'scaling': 'strong'}
    if mat_ms is not None:
        gb = U * a.items * 4 / 1e9
        line['materialise_ms'] = mat_ms
        line['roofline_materialise'] = {'bound': 'hbm', 'achieved': gb / (mat_ms / 1e3), 'peak': peaks['hbm_gbs'],
                                        'unit': 'GB/s', 'frac': gb / (mat_ms / 1e3) / peaks['hbm_gbs'],
                                        'note': 'algorithmic bytes = 4 B per score written'}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
