"""Times the HOST side of the path at a preset's shape (no GPU needed): dataset load, test / validation set build
(1000 negatives per distinct user), and per epoch the shuffle (§8a-2), the training negatives (§8a-3)
and the batch assembly (§8a-4).  Usage: python tools/host_pipeline_bench.py [--preset electronics] [--epochs 2]
Writes nothing outside a temporary directory; prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

from dccf_b200 import synth                                      # noqa: E402
from dccf_b200.data_loaders.DataLoader import DataLoader         # noqa: E402
from dccf_b200.data_processor.DataProcessor import DataProcessor # noqa: E402


class _ModelFlags(object):
    """What DataProcessor reads from the model class (RecModel.py:10-13)."""
    append_id, include_id = True, False
    include_user_features = include_item_features = include_context_features = False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--preset', default='electronics', choices=sorted(synth.PRESETS))
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--test_neg_n', type=int, default=1000)
    ap.add_argument('--batch_size', type=int, default=128)
    ap.add_argument('--eval_batch_size', type=int, default=16384)
    ap.add_argument('--python-sampler', action='store_true', help='the draw-for-draw Python loop instead of C++')
    a = ap.parse_args()
    U, I, per = synth.PRESETS[a.preset]
    out = {'preset': a.preset, 'users': U, 'items': I}
    with tempfile.TemporaryDirectory() as tmp:
        t = time.perf_counter()
        d = os.path.join(tmp, a.preset)
        os.makedirs(d)
        inter = synth.make_interactions(U, I, per)
        for name in ('train', 'validation', 'test'):
            u, i, l, tt = inter[name]
            np.savetxt(os.path.join(d, '%s.%s.csv' % (a.preset, name)), np.stack([u, i, l, tt], axis=1), fmt='%d',
                       delimiter=',')
        out['synth_write_s'] = time.perf_counter() - t

        def stage(name, fn):
            t0 = time.perf_counter()
            r = fn()
            out[name] = round(time.perf_counter() - t0, 3)
            return r

        np.random.seed(2019)
        loader = stage('load_first_s', lambda: DataLoader(tmp, a.preset, sep=','))
        loader = stage('load_cached_s', lambda: DataLoader(tmp, a.preset, sep=','))
        stage('drop_neg_s', loader.drop_neg)
        dp = stage('dp_init_s', lambda: DataProcessor(loader, _ModelFlags, rank=1, test_neg_n=a.test_neg_n))
        dp.use_native_sampler = not a.python_sampler
        test = stage('test_data_s', dp.get_test_data)
        out['test_rows'] = int(len(test['Y']))
        stage('test_batches_s', lambda: dp.prepare_batches(test, a.eval_batch_size, train=False))
        val = stage('validation_data_s', dp.get_validation_data)
        stage('validation_batches_s', lambda: dp.prepare_batches(val, a.eval_batch_size, train=False))
        for e in range(a.epochs):
            data = stage('epoch%d_shuffle_s' % e, lambda: dp.get_train_data(e))
            b = stage('epoch%d_batches_s' % e, lambda: dp.prepare_batches(data, a.batch_size, train=True))
            out['train_rows'] = int(len(data['Y']))
            out['train_batches'] = len(b)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
