"""Model base class with the protocol of the reference's src/models/BaseModel.py:15-248: feature flags,
`parse_model_args`, `evaluate_method`, `init_paras`, `l2`, `forward` (BPR / MSE), `save_model/load_model`.

`evaluate_method` keeps the reference signature (numpy predictions + data dict + metric names) but
the ranking metrics are computed by the dccf_rank_eval CUDA kernel, not by pandas."""
import logging
import os

import numpy as np
import torch
import torch.nn.functional as F

from .. import kernels


def group_candidates(uid):
    """Rows grouped by user: (sorted distinct users, cand_rows int32 [n], user_off int64 [n_users+1]).
    Stable, so rows of one user keep data order."""
    uid = np.asarray(uid)
    order = np.argsort(uid, kind='stable')
    users, starts = np.unique(uid[order], return_index=True)
    off = np.concatenate([starts, [len(uid)]]).astype(np.int64)
    return users, order.astype(np.int32), off


def candidate_layout(uid):
    """(cand_rows or None, user_off) for the ranker.  When every user's rows already form ONE contiguous run of the data
    (an evaluation set written user by user) the ranker reads scores / labels in place: cand_rows is None and user_off
    holds the run boundaries (users in data order).  Otherwise (the reference's layout: all positives, then the blocks
    of negatives, src/data_processor/DataProcessor.py:92-111) rows are grouped through an index array."""
    uid = np.asarray(uid)
    n = len(uid)
    if n == 0:
        return None, np.zeros(1, dtype=np.int64)
    starts = np.concatenate([[0], np.flatnonzero(uid[1:] != uid[:-1]) + 1])
    if len(np.unique(uid[starts])) == len(starts):
        return None, np.concatenate([starts, [n]]).astype(np.int64)
    _, rows, off = group_candidates(uid)
    return rows, off


def rank_metrics_device(scores, labels, iids, cand_rows, user_off, k, want_topk=False):
    """Launch dccf_rank_eval; all arguments are CUDA tensors (cand_rows None: rows already grouped by user).
    Returns per-user metrics [n_users,5] f64 (ndcg, hit, precision, recall, f1 at k) and optionally the top-k
    item ids."""
    n_users = user_off.shape[0] - 1
    out = torch.empty((n_users, 5), dtype=torch.float64, device=scores.device)
    topk = torch.empty((n_users, k), dtype=torch.int64, device=scores.device) if want_topk else None
    kernels.rank_eval(scores, labels, iids, cand_rows, user_off, k, out, out_topk_iid=topk)
    return (out, topk) if want_topk else out


def rank_sums_device(scores, labels, iids, cand_rows, user_off, ks):
    """Sums over users of (ndcg, hit, precision, recall, f1) at every k of `ks` -> CUDA tensor [len(ks), 5] f64, rows in
    the order of `ks`.  One launch of dccf_rank_eval_multi per group of up to four values of k <= 16 (a metric list such
    as ndcg@5,recall@5,precision@5 is ONE launch; the reference loops over the users once per metric,
    src/models/BaseModel.py:90-126); larger k go through dccf_rank_eval one at a time."""
    ks = [int(k) for k in ks]
    small = sorted(set(k for k in ks if k <= kernels.RANK_STREAM_MAX_K))
    if ks == small and len(ks) <= kernels.RANK_MAX_NK:          # the usual case: one launch, its output is the result
        sums = torch.empty((len(ks), 5), dtype=torch.float64, device=scores.device)
        kernels.rank_eval_multi(scores, labels, iids, cand_rows, user_off, ks, out_sums=sums)
        return sums
    out = torch.empty((len(ks), 5), dtype=torch.float64, device=scores.device)
    where = {}
    for a in range(0, len(small), kernels.RANK_MAX_NK):
        group = small[a:a + kernels.RANK_MAX_NK]
        sums = torch.empty((len(group), 5), dtype=torch.float64, device=scores.device)
        kernels.rank_eval_multi(scores, labels, iids, cand_rows, user_off, group, out_sums=sums)
        for j, k in enumerate(group):
            where[k] = sums[j]
    for k in set(ks) - set(small):
        where[k] = rank_metrics_device(scores, labels, iids, cand_rows, user_off, k).sum(dim=0)
    for j, k in enumerate(ks):
        out[j] = where[k]
    return out


METRIC_COLUMN = {'ndcg': 0, 'hit': 1, 'precision': 2, 'recall': 3, 'f1': 4}

# device-resident grouping (rows by user, labels, item ids) per data dict, keyed by the identity of its
# 'uid' array: validation/test dicts are built once and evaluated every epoch
_RANK_CTX = {}


def rank_context(data, dev):
    uid = data['uid']
    key = id(uid)
    hit = _RANK_CTX.get(key)
    if hit is not None and hit[0]() is uid and hit[1]['dev'] == dev and hit[1]['n'] == len(uid):
        return hit[1]
    import weakref
    rows, off = candidate_layout(uid)
    ctx = {'dev': dev, 'n': len(uid),
           'labels': torch.from_numpy(np.ascontiguousarray(data['Y'], dtype=np.float32)).to(dev),
           'iids': torch.from_numpy(np.ascontiguousarray(data['iid'], dtype=np.int64)).to(dev),
           'rows': None if rows is None else torch.from_numpy(rows).to(dev), 'off': torch.from_numpy(off).to(dev),
           'n_users': len(off) - 1}
    try:
        _RANK_CTX[key] = (weakref.ref(uid, lambda _r, k=key: _RANK_CTX.pop(k, None)), ctx)
    except TypeError:
        pass
    return ctx


def _rank_metric_sums(p, data, metrics):
    """{(name, k): sum over users} for every '<name>@k' metric of the list, plus the user count — one ranker launch
    and one device -> host copy for the whole list."""
    wanted = []
    for metric in metrics:
        if '@' in metric:
            name, k = metric.split('@')
            if name in METRIC_COLUMN:
                wanted.append((name, int(k)))
    if not wanted:
        return {}, 0
    if not torch.cuda.is_available():
        raise RuntimeError('ranking metrics run on the GPU (dccf_rank_eval); no CUDA device is visible')
    dev = p.device if (torch.is_tensor(p) and p.is_cuda) else torch.device('cuda', torch.cuda.current_device())
    ctx = rank_context(data, dev)
    scores = p.to(dev, torch.float32).contiguous() if torch.is_tensor(p) else \
        torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32)).to(dev)
    ks = sorted(set(k for _, k in wanted))
    sums = rank_sums_device(scores, ctx['labels'], ctx['iids'], ctx['rows'], ctx['off'], ks).cpu().numpy()
    return {(name, k): float(sums[ks.index(k), METRIC_COLUMN[name]]) for name, k in wanted}, ctx['n_users']


class BaseModel(torch.nn.Module):
    # how DataProcessor.format_data_dict lays out 'X' (BaseModel.py:36-40)
    append_id = False
    include_id = True
    include_user_features = True
    include_item_features = True
    include_context_features = False

    @staticmethod
    def parse_model_args(parser, model_name='BaseModel'):
        parser.add_argument('--model_path', type=str, default='../model/%s/%s.pt' % (model_name, model_name),
                            help='Model save path.')
        return parser

    @staticmethod
    def evaluate_method(p, data, metrics):
        """Metric values for predictions `p` aligned with `data` (BaseModel.py:56-128).
        'rmse'/'mae' and the sklearn classification metrics are host arithmetic; every '<name>@k'
        metric (ndcg, hit, precision, recall, f1) comes from the GPU ranker, averaged over users."""
        l = data['Y']
        evaluations = []
        rank_sums, n_users = _rank_metric_sums(p, data, metrics)
        for metric in metrics:
            if metric == 'rmse':
                d = np.asarray(l, dtype=np.float64) - np.asarray(p, dtype=np.float64)
                evaluations.append(float(np.sqrt(np.mean(d * d))))
            elif metric == 'mae':
                d = np.asarray(l, dtype=np.float64) - np.asarray(p, dtype=np.float64)
                evaluations.append(float(np.mean(np.abs(d))))
            elif metric in ('auc', 'f1', 'accuracy', 'precision', 'recall'):
                import sklearn.metrics as skm
                fn = {'auc': skm.roc_auc_score, 'f1': skm.f1_score, 'accuracy': skm.accuracy_score,
                      'precision': skm.precision_score, 'recall': skm.recall_score}[metric]
                evaluations.append(fn(l, p))
            else:
                name, k = metric.split('@')
                if name not in METRIC_COLUMN:
                    continue
                evaluations.append(rank_sums[(name, int(k))] / max(1, n_users))
        return evaluations

    @staticmethod
    def evaluate_sums(p, data, metrics):
        """(sums, counts) per metric over the rows of `data`, so that partial results of user shards can be
        added across ranks: rank metrics sum over users, mae sums |err| and rmse sums err^2 over rows."""
        sums, counts = [], []
        rank_sums, n_users = _rank_metric_sums(p, data, metrics)
        for metric in metrics:
            if metric in ('rmse', 'mae'):
                pl = p.detach().cpu().numpy() if torch.is_tensor(p) else np.asarray(p)
                d = np.asarray(data['Y'], dtype=np.float64) - pl.astype(np.float64)
                sums.append(float(np.sum(d * d) if metric == 'rmse' else np.sum(np.abs(d))))
                counts.append(float(len(d)))
                continue
            if '@' not in metric or metric.split('@')[0] not in METRIC_COLUMN:
                raise ValueError('metric %r cannot be summed over user shards (use rmse, mae or <name>@k)' % metric)
            name, k = metric.split('@')
            sums.append(rank_sums[(name, int(k))])
            counts.append(float(n_users))
        return sums, counts

    @staticmethod
    def init_paras(m):
        """N(0, 0.01) for Linear weights/biases and Embedding weights (BaseModel.py:131-141); applied
        through `model.apply` in module order, on the torch CPU generator (src/main.py:150)."""
        if type(m) == torch.nn.Linear:
            torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)
            if m.bias is not None:
                torch.nn.init.normal_(m.bias, mean=0.0, std=0.01)
        elif type(m) == torch.nn.Embedding:
            torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)

    def __init__(self, label_min, label_max, feature_num, random_seed=2018, model_path='../model/Model/Model.pt'):
        super(BaseModel, self).__init__()
        self.label_min = label_min
        self.label_max = label_max
        self.feature_num = feature_num
        self.random_seed = random_seed
        torch.manual_seed(self.random_seed)           # BaseModel.py:150-151
        if torch.cuda.is_available():
            torch.cuda.manual_seed(self.random_seed)
        self.model_path = model_path
        self._init_weights()
        logging.debug(list(self.parameters()))
        self.total_parameters = self.count_variables()
        logging.info('# of params: %d' % self.total_parameters)
        self.optimizer = None                         # set by the runner (BaseModel.py:161)

    def _init_weights(self):
        self.x_bn = torch.nn.BatchNorm1d(self.feature_num)
        self.prediction = torch.nn.Linear(self.feature_num, 1)

    def count_variables(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def l2(self):
        """Sum of squares of ALL parameters (BaseModel.py:179-187)."""
        total = 0
        for p in self.parameters():
            total = total + (p ** 2).sum()
        return total

    def predict(self, feed_dict):
        x = self.x_bn(feed_dict['X'].float())
        x = torch.nn.Dropout(p=feed_dict['dropout'])(x)
        prediction = F.relu(self.prediction(x)).view([-1])
        return {'prediction': prediction, 'check': []}

    def forward(self, feed_dict):
        """predict + loss: BPR over [positives ; negatives] for rank 1, MSE otherwise (BaseModel.py:203-219)."""
        out_dict = self.predict(feed_dict)
        if feed_dict['rank'] == 1:
            b = int(feed_dict['Y'].shape[0] / 2)
            pos, neg = out_dict['prediction'][:b], out_dict['prediction'][b:]
            loss = -(pos - neg).sigmoid().log().sum()
        else:
            loss = torch.nn.MSELoss()(out_dict['prediction'], feed_dict['Y'])
        out_dict['loss'] = loss
        return out_dict

    def lrp(self):
        pass

    def save_model(self, model_path=None):
        """state_dict only, same keys as the reference so .pt files interchange (BaseModel.py:224-236)."""
        model_path = model_path or self.model_path
        multi = torch.distributed.is_available() and torch.distributed.is_initialized()
        if not multi or torch.distributed.get_rank() == 0:     # replicas are identical: rank 0 writes for all
            dir_path = os.path.dirname(model_path)
            if dir_path and not os.path.exists(dir_path):
                os.makedirs(dir_path)
            torch.save(self.state_dict(), model_path)
        if multi:
            torch.distributed.barrier()
        logging.info('Save model to ' + model_path)

    def load_model(self, model_path=None):
        model_path = model_path or self.model_path
        multi = torch.distributed.is_available() and torch.distributed.is_initialized()
        # (under torchrun the file was written from rank 0's GPU: stage through the host instead of that device)
        self.load_state_dict(torch.load(model_path, map_location='cpu') if multi else torch.load(model_path))
        self.eval()
        logging.info('Load model from ' + model_path)
