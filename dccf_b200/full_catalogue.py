"""Full-catalogue scoring on the tcgen05 tensor cores (dccf_full_scores) — the one dense user x item GEMM of
the path (BASELINE.json config 4).

* `ipsmf_exposure(factors)` — the IPSBiasedMF prediction of every (user, item) pair
  (src/models/IPSBiasedMF.py:37-57): `(<p_u,q_i> + b_u + b_i + g) / max(propensity_i, M)`.  README.md:27-29 of the
  reference: this matrix is what DCCF loads as `<dataset>.ips_expo_prob.npy`.
* `ipsmf_topk(factors, k)` — the same scores, per-user top-k only (the matrix is never written).
* `dccf_catalogue_topk(model, k)` — deterministic DCCF scores of the whole catalogue (std 0, no confounder
  items): `score[u,i] = <E_user[u], relu(W_i·E_item[i] + W_f·Feat[i] + b)>`.
"""
import torch

from . import kernels


def _col_scale(factors):
    prop = factors['propensity']
    return 1.0 / torch.clamp(prop, min=float(factors['mf_min_propensity']))


def ipsmf_exposure(factors):
    """[U,I] float32 exposure matrix from IPSBiasedMF factors (dict of CUDA tensors, keys as in
    dccf_b200.synth.make_ipsmf_factors)."""
    out, _, _ = kernels.full_scores(factors['mf_user'], factors['mf_item'], factors['mf_user_bias'],
                                    factors['mf_item_bias'], _col_scale(factors), float(factors['mf_global_bias']),
                                    materialise=True, k=0)
    return out


def ipsmf_topk(factors, k):
    _, s, i = kernels.full_scores(factors['mf_user'], factors['mf_item'], factors['mf_user_bias'],
                                  factors['mf_item_bias'], _col_scale(factors), float(factors['mf_global_bias']),
                                  materialise=False, k=k)
    return s, i


def dccf_catalogue_topk(model, k, users=None):
    """Top-k items per user over the WHOLE catalogue with the deterministic DCCF predictor."""
    if model.sample_num != 0 or model.std != 0:
        raise ValueError('full-catalogue DCCF scoring is a GEMM only without confounder samples and feature noise '
                         '(sample_num == 0, std == 0); got sample_num=%d std=%g' % (model.sample_num, model.std))
    model._check_ready()
    t = model._tc_tables()
    hidden = torch.relu(t['PI'] + t['PF']).contiguous()
    A = model.uid_embeddings.weight.data if users is None else model.uid_embeddings.weight.data[users].contiguous()
    _, s, i = kernels.full_scores(A, hidden, materialise=False, k=k)
    return s, i
