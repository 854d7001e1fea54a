"""Host (numpy) versions of the ranking helpers the reference exposes in src/utils/rank_metrics.py,
kept for API compatibility (`from utils.rank_metrics import *`).  The evaluation hot path does NOT use
these: it runs dccf_rank_eval on the GPU (dccf_b200/csrc/rank_eval.cu)."""
import numpy as np


def precision_at_k(r, k):
    """Fraction of non-zero relevances among the first k (src/utils/rank_metrics.py:61-87)."""
    assert k >= 1
    r = np.asarray(r)[:k] != 0
    if r.size != k:
        raise ValueError('Relevance score length < k')
    return np.mean(r)


def dcg_at_k(r, k, method=0):
    """method 0: r0 + sum_{j>=1} r_j/log2(j+1); method 1: sum r_j/log2(j+2)
    (src/utils/rank_metrics.py:130-167)."""
    r = np.asarray(r, dtype=np.float64)[:k]
    if r.size:
        if method == 0:
            return r[0] + np.sum(r[1:] / np.log2(np.arange(2, r.size + 1)))
        if method == 1:
            return np.sum(r / np.log2(np.arange(2, r.size + 2)))
        raise ValueError('method must be 0 or 1.')
    return 0.


def ndcg_at_k(r, k, method=0):
    """DCG normalised by the DCG of the ideal ordering (src/utils/rank_metrics.py:170-201)."""
    dcg_max = dcg_at_k(sorted(r, reverse=True), k, method)
    if not dcg_max:
        return 0.
    return dcg_at_k(r, k, method) / dcg_max
