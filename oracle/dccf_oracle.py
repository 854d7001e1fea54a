"""TEST INFRASTRUCTURE — not part of the product path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

CPU restatement (numpy, float32 or float64) of the reference's DCCF hot path, function by function.
Parity status: PINNED — `tests/test_oracle_golden.py` checks every function below against fixtures
produced by the unmodified reference itself (`oracle/make_golden.py`, run in the build container
through `oracle/ref_harness.py`) and against the known-answer values in the reference's own
docstrings (src/utils/rank_metrics.py:64-70, 136-148, 176-187).

Notation (src/models/DCCF.py): P pairs, S = --sample-num, Z = S+1 slots, A = --attribute-num,
R = Z*A rows per pair, row r = (p*Z + z)*A + a, D = 64, F = feature width.
"""
import numpy as np


# ------------------------------------------------------------------------------------------------
# DCCF.predict  (src/models/DCCF.py:66-107)
# ------------------------------------------------------------------------------------------------
def slot_items(X, sample_item):
    """items[p, 0] = the true item, items[p, 1:] = sampled confounders  (DCCF.py:74)."""
    X = np.asarray(X)
    if sample_item is None or np.asarray(sample_item).size == 0:
        return X[:, 1:2].astype(np.int64)
    return np.concatenate([X[:, 1:2], np.asarray(sample_item)], axis=1).astype(np.int64)


def exposure_values(expo, u, items):
    """expo_prob[u, item]  (DCCF.py:98) — dense matrix, or on the fly from IPSBiasedMF factors
    (src/models/IPSBiasedMF.py:37-57): (<p_u,q_i> + b_u + b_i + g) / max(propensity_i, M)."""
    if isinstance(expo, dict):
        pu = expo['mf_user'][u]                      # [P, D]
        qi = expo['mf_item'][items]                  # [P, Z, D]
        dot = np.einsum('pd,pzd->pz', pu, qi)
        pred = dot + expo['mf_user_bias'][u][:, None] + expo['mf_item_bias'][items] + expo['mf_global_bias']
        prop = np.maximum(expo['propensity'][items], expo['mf_min_propensity'])
        return pred / prop
    return np.asarray(expo)[u[:, None], items]


def predict(params, X, sample_item, noise, mask, A, dtype=np.float32, expo=None):
    """Returns dict(pred[P], h[N,D] post-dropout activations, w[P,Z] softmax exposure weights, s[P,Z,A]).

    params: dict with E_user [U,D], E_item [I,D], W [D, D+F], b [D], Feat [I,F], expo [U,I] (or `expo=` dict)
    noise: [N,F] (already scaled by std) or None; mask: [N,D] dropout multipliers or None.
    """
    X = np.asarray(X)
    P = X.shape[0]
    items = slot_items(X, sample_item)               # [P, Z]
    Z = items.shape[1]
    E_user = params['E_user'].astype(dtype)
    E_item = params['E_item'].astype(dtype)
    W = params['W'].astype(dtype)
    b = params['b'].astype(dtype)
    Feat = params['Feat'].astype(dtype)
    D = E_user.shape[1]
    F = Feat.shape[1]
    u = X[:, 0].astype(np.int64)
    fi = X[:, 1].astype(np.int64)
    N = P * Z * A

    item_emb = np.broadcast_to(E_item[items][:, :, None, :], (P, Z, A, D)).reshape(N, D)      # DCCF.py:85
    feat = np.broadcast_to(Feat[fi][:, None, None, :], (P, Z, A, F)).reshape(N, F)            # DCCF.py:86
    if noise is not None:
        feat = feat + np.asarray(noise).astype(dtype).reshape(N, F)                           # DCCF.py:87
    x = np.concatenate([item_emb, feat], axis=1)                                              # DCCF.py:89
    pre = x @ W.T + b                                                                         # DCCF.py:92
    h = np.maximum(pre, 0)                                                                    # DCCF.py:93
    if mask is not None:
        h = h * np.asarray(mask).astype(dtype).reshape(N, D)                                  # DCCF.py:94
    user_emb = np.broadcast_to(E_user[u][:, None, None, :], (P, Z, A, D)).reshape(N, D)       # DCCF.py:84
    s = (user_emb * h).sum(axis=1).reshape(P, Z, A)                                           # DCCF.py:96
    ev = exposure_values(params['expo'] if expo is None else expo, u, items).astype(dtype)    # DCCF.py:98
    ev = ev - ev.max(axis=1, keepdims=True)
    e = np.exp(ev)
    w = e / e.sum(axis=1, keepdims=True)
    pred = (w[:, :, None] * s).sum(axis=1).mean(axis=1)                                       # DCCF.py:100
    return {'pred': pred.astype(dtype), 'h': h, 'w': w, 's': s, 'x': x, 'pre': pre, 'items': items}


# ------------------------------------------------------------------------------------------------
# DCCF.forward loss  (src/models/DCCF.py:109-127)
# ------------------------------------------------------------------------------------------------
def loss_bpr(pred):
    b = pred.shape[0] // 2
    d = pred[:b].astype(np.float64) - pred[b:2 * b].astype(np.float64)
    return float(np.sum(np.logaddexp(0.0, -d)))      # -log sigmoid(d)


def loss_mse(pred, Y):
    d = pred.astype(np.float64) - np.asarray(Y, dtype=np.float64)
    return float(np.mean(d * d))


# ------------------------------------------------------------------------------------------------
# backward of loss wrt the four parameter tensors (what autograd does at BaseRunner.py:183),
# WITHOUT the l2 term (that one is part of `adam_step`)
# ------------------------------------------------------------------------------------------------
def backward(params, X, sample_item, noise, mask, A, fwd, loss_mode=0, Y=None, dtype=np.float64):
    X = np.asarray(X)
    P = X.shape[0]
    items = fwd['items']
    Z = items.shape[1]
    E_user = params['E_user'].astype(dtype)
    W = params['W'].astype(dtype)
    D = E_user.shape[1]
    u = X[:, 0].astype(np.int64)
    pred = fwd['pred'].astype(dtype)
    if loss_mode == 0:
        b = P // 2
        sg = 1.0 / (1.0 + np.exp(-(pred[:b] - pred[b:2 * b])))
        dpred = np.zeros(P, dtype=dtype)
        dpred[:b] = -(1.0 - sg)
        dpred[b:2 * b] = (1.0 - sg)
    else:
        dpred = 2.0 * (pred - np.asarray(Y, dtype=dtype)) / P
    w = fwd['w'].astype(dtype)
    ds = (dpred[:, None] * w / A)[:, :, None] * np.ones((1, 1, A), dtype=dtype)               # [P,Z,A]
    h = fwd['h'].astype(dtype).reshape(P, Z, A, D)
    gate = (fwd['pre'].reshape(P, Z, A, D) > 0).astype(dtype)
    if mask is not None:
        gate = gate * np.asarray(mask).astype(dtype).reshape(P, Z, A, D)
    gu_rows = (ds[..., None] * h).sum(axis=(1, 2))                                            # [P,D]
    dpre = ds[..., None] * E_user[u][:, None, None, :] * gate                                 # [P,Z,A,D]
    dpre2 = dpre.reshape(P * Z * A, D)
    gb = dpre2.sum(axis=0)
    gW = dpre2.T @ fwd['x'].astype(dtype)                                                     # [D, D+F]
    gi_rows = dpre.sum(axis=2) @ W[:, :D]                                                     # [P,Z,D]
    gE_user = np.zeros_like(params['E_user'], dtype=dtype)
    gE_item = np.zeros_like(params['E_item'], dtype=dtype)
    np.add.at(gE_user, u, gu_rows)
    np.add.at(gE_item, items.reshape(-1), gi_rows.reshape(P * Z, D))
    return {'E_user': gE_user, 'E_item': gE_item, 'W': gW, 'b': gb,
            'gu_rows': gu_rows, 'gi_rows': gi_rows.reshape(P * Z, D), 'dpred': dpred}


# ------------------------------------------------------------------------------------------------
# loss += l2*sum p^2 ; clip_grad_value_(50) ; Adam(weight_decay=l2).step()
# (src/runners/BaseRunner.py:181,185,100,187; torch 2.11 optim/adam.py non-capturable branch)
# ------------------------------------------------------------------------------------------------
def adam_step(p, g_data, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, l2=1e-4, weight_decay=1e-4, clip=50.0,
              dtype=np.float32):
    f = dtype
    p = p.astype(f)
    g = g_data.astype(f) + f(2.0 * l2) * p
    if clip > 0:
        g = np.clip(g, f(-clip), f(clip))
    g = g + f(weight_decay) * p
    m = m.astype(f) + f(1.0 - beta1) * (g - m.astype(f))
    v = v.astype(f) * f(beta2) + f(1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    step_size = f(-(lr / bc1))
    denom = np.sqrt(v) / f(bc2 ** 0.5) + f(eps)
    p = p + step_size * (m / denom)
    return p.astype(f), m.astype(f), v.astype(f)


def train_step(params, state, t, X, sample_item, noise, mask, A, hp, loss_mode=0, Y=None, grad_dtype=np.float64,
               dtype=np.float32):
    """One BaseRunner.fit iteration (BaseRunner.py:175-188).  `state` holds m/v per parameter.
    Returns (new_params, new_state, loss, pred)."""
    fwd = predict(params, X, sample_item, noise, mask, A, dtype=grad_dtype)
    loss = loss_bpr(fwd['pred']) if loss_mode == 0 else loss_mse(fwd['pred'], Y)
    grads = backward(params, X, sample_item, noise, mask, A, fwd, loss_mode=loss_mode, Y=Y, dtype=grad_dtype)
    new_params = dict(params)
    new_state = {}
    for k in ('E_user', 'E_item', 'W', 'b'):
        p, m, v = adam_step(params[k], grads[k], state[k]['m'], state[k]['v'], t, dtype=dtype, **hp)
        new_params[k] = p
        new_state[k] = {'m': m, 'v': v}
    return new_params, new_state, loss, fwd['pred']


# ------------------------------------------------------------------------------------------------
# ranking metrics  (src/models/BaseModel.py:82-126, src/utils/rank_metrics.py:61-87,130-201)
# ------------------------------------------------------------------------------------------------
def dcg_at_k(r, k):
    """method=1 DCG: sum r_j / log2(j + 2)  (rank_metrics.py:159-164)."""
    r = np.asarray(r, dtype=np.float64)[:k]
    if r.size:
        return float(np.sum(r / np.log2(np.arange(2, r.size + 2))))
    return 0.0


def ndcg_at_k(r, k):
    dcg_max = dcg_at_k(sorted(r, reverse=True), k)                                            # rank_metrics.py:198
    if not dcg_max:
        return 0.0
    return dcg_at_k(r, k) / dcg_max


def rank_users(scores, uid, labels, iids, k):
    """Per-user top-k and metrics with the total order (score desc, iid asc, row asc); NaN last.
    Returns (user_ids sorted ascending, topk_iid [n_users,k] (-1 padded), topk_row, metrics [n_users,5]
    = ndcg@k, hit@k, precision@k, recall@k, f1@k)."""
    scores = np.asarray(scores, dtype=np.float32)
    uid = np.asarray(uid)
    labels = np.asarray(labels, dtype=np.float32)
    iids = np.asarray(iids, dtype=np.int64)
    rows = np.arange(len(scores))
    key_s = np.where(np.isnan(scores), -np.inf, scores)
    order = np.lexsort((rows, iids, -key_s.astype(np.float64), uid))
    users, starts = np.unique(uid[order], return_index=True)
    bounds = list(starts) + [len(order)]
    n_users = len(users)
    topk_iid = -np.ones((n_users, k), dtype=np.int64)
    topk_row = -np.ones((n_users, k), dtype=np.int32)
    metrics = np.zeros((n_users, 5), dtype=np.float64)
    for g in range(n_users):
        idx = order[bounds[g]:bounds[g + 1]]
        l = labels[idx].astype(np.float64)
        top = idx[:k]
        topk_iid[g, :len(top)] = iids[top]
        topk_row[g, :len(top)] = top
        lt = l[:k]
        hits = float(np.sum(lt))
        tot = float(np.sum(l))
        metrics[g, 0] = ndcg_at_k(list(l), k)
        metrics[g, 1] = 1.0 if hits > 0 else 0.0
        metrics[g, 2] = float(np.sum(lt != 0)) / k
        metrics[g, 3] = hits / tot if tot != 0 else np.nan
        metrics[g, 4] = 2.0 * hits / (k + tot)
    return users, topk_iid, topk_row, metrics


METRIC_COLUMN = {'ndcg': 0, 'hit': 1, 'precision': 2, 'recall': 3, 'f1': 4}


def evaluate_method(p, data, metrics):
    """BaseModel.evaluate_method for the '<name>@k' metrics (BaseModel.py:82-126): mean over users."""
    out = []
    for metric in metrics:
        name, k = metric.split('@')
        _, _, _, m = rank_users(p, data['uid'], data['Y'], data['iid'], int(k))
        out.append(float(np.average(m[:, METRIC_COLUMN[name]])))
    return out


# ------------------------------------------------------------------------------------------------
# numpy-legacy / torch-CPU random primitives used for index parity
# ------------------------------------------------------------------------------------------------
class MT19937:
    """The generator behind both `np.random.seed(s)` (legacy) and torch's CPU generator."""

    def __init__(self, seed):
        self.mt = np.zeros(624, dtype=np.uint64)
        self.mt[0] = seed & 0xffffffff
        for i in range(1, 624):
            prev = int(self.mt[i - 1])
            self.mt[i] = (1812433253 * (prev ^ (prev >> 30)) + i) & 0xffffffff
        self.idx = 624

    def _twist(self):
        mt = [int(x) for x in self.mt]
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7fffffff)
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
        self.mt = np.array(mt, dtype=np.uint64)
        self.idx = 0

    def next_u32(self):
        if self.idx >= 624:
            self._twist()
        y = int(self.mt[self.idx])
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return y & 0xffffffff

    def np_randint(self, n):
        """np.random.randint(n): masked rejection (legacy bounded integers)."""
        rng = n - 1
        if rng == 0:
            return 0
        mask = rng
        for s in (1, 2, 4, 8, 16):
            mask |= mask >> s
        while True:
            v = self.next_u32() & mask
            if v <= rng:
                return v

    def torch_randint(self, n, count):
        """torch.randint(n, (count,)) on the CPU generator: next_u32 % n, sequential."""
        return np.array([self.next_u32() % n for _ in range(count)], dtype=np.int64)
