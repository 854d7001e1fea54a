"""Host helpers with the reference's names and behaviour (src/utils/utils.py)."""
import logging
import os

import numpy as np
import torch

LOWER_METRIC_LIST = ['rmse', 'mae']


def parse_global_args(parser):
    """Global flags and defaults of src/utils/utils.py:10-28."""
    parser.add_argument('--gpu', type=str, default='0', help='Set CUDA_VISIBLE_DEVICES')
    parser.add_argument('--verbose', type=int, default=logging.INFO, help='Logging Level, 0, 10, ..., 50')
    parser.add_argument('--log_file', type=str, default='../log/log.txt', help='Logging file path')
    parser.add_argument('--result_file', type=str, default='../result/result.npy', help='Result file path')
    parser.add_argument('--random_seed', type=int, default=2019, help='Random seed of numpy and torch.')
    parser.add_argument('--train', type=int, default=1, help='To train the model or not.')
    return parser


def format_metric(metric):
    """'%.4f' for floats, '%d' for ints, comma-joined (src/utils/utils.py:62-79)."""
    if not isinstance(metric, (tuple, list)):
        metric = [metric]
    parts = []
    for m in metric:
        if isinstance(m, (float, np.floating)):
            parts.append('%.4f' % m)
        elif isinstance(m, (int, np.integer)):
            parts.append('%d' % m)
    return ','.join(parts)


def shuffle_in_unison_scary(data):
    """Shuffle every array of the dict with the SAME permutation by replaying one numpy RNG state per
    array; the global stream advances by one shuffle (src/utils/utils.py:82-92)."""
    arrays = [data[key] for key in data]
    lengths = {len(a) for a in arrays}
    distinct = all(not np.may_share_memory(a, b) for k, a in enumerate(arrays) for b in arrays[k + 1:]
                   if isinstance(a, np.ndarray) and isinstance(b, np.ndarray))
    if len(lengths) == 1 and distinct and all(isinstance(a, np.ndarray) for a in arrays):
        # Same result, one Fisher-Yates pass instead of one per array: np.random.shuffle draws j in [0, i] for
        # i = n-1 .. 1 whatever the array holds (1-D or 2-D), so shuffling arange(n) from the current state yields the
        # permutation every array would undergo and leaves the generator exactly where the last replay would have.
        # Arrays are overwritten in place (their identity is part of the reference's behaviour: the train dict handed
        # out by get_train_data(-1) is the one later epochs shuffle).
        perm = np.arange(lengths.pop())
        np.random.shuffle(perm)
        for a in arrays:
            a[...] = a[perm]
        return data
    state = np.random.get_state()
    for a in arrays:
        np.random.set_state(state)
        np.random.shuffle(a)
    return data


def best_result(metric, results_list):
    """min for rmse/mae, max otherwise; lists compare lexicographically (src/utils/utils.py:95-106)."""
    if isinstance(metric, (list, tuple)):
        metric = metric[0]
    return min(results_list) if metric in LOWER_METRIC_LIST else max(results_list)


def strictly_increasing(l):
    return all(a < b for a, b in zip(l, l[1:]))


def strictly_decreasing(l):
    return all(a > b for a, b in zip(l, l[1:]))


def non_increasing(l):
    return all(a >= b for a, b in zip(l, l[1:]))


def non_decreasing(l):
    return all(a <= b for a, b in zip(l, l[1:]))


def monotonic(l):
    return non_increasing(l) or non_decreasing(l)


def numpy_to_torch(d):
    """numpy -> torch, on the GPU when one is visible (src/utils/utils.py:154-163)."""
    t = torch.from_numpy(d)
    if torch.cuda.device_count() > 0:
        t = t.cuda()
    return t


def tensor_to_gpu(t):
    return t.cuda() if torch.cuda.device_count() > 0 else t


def check_dir_and_mkdir(path):
    """Create the directory of `path` (or `path` itself when it names a directory)
    (src/utils/utils.py:171-178)."""
    if os.path.basename(path).find('.') == -1 or path.endswith('/'):
        dirname = path
    else:
        dirname = os.path.dirname(path)
    if dirname and not os.path.exists(dirname):
        print('make dirs:', dirname)
        os.makedirs(dirname, exist_ok=True)         # (several ranks may get here at once under torchrun)
