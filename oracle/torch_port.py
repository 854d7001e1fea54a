"""TEST / BASELINE INFRASTRUCTURE — not part of the product path.

Eager-PyTorch CPU port of the reference's DCCF hot path, op for op (same ATen calls in the same order
as src/models/DCCF.py:66-127 and src/runners/BaseRunner.py:175-188), used ONLY by bench.py as the
`cpu_baseline` / `--impl reference` arm: the reference itself is Python under /root/reference and cannot
travel to the GPU box, and its DCCF class hard-codes CUDA (DCCF.py:55,64,72,87).  Validated against the
reference fixtures in tests/test_oracle_golden.py::test_torch_port_matches_reference.

`to_device('cuda')` moves the parameters AND the two plain-attribute tables, after which the same op sequence runs
eagerly on the GPU exactly as the reference does on its own hardware (confounders on the CPU generator then copied,
noise / dropout on the CUDA generator): bench.py's `gpu_eager_reference` object — SURVEY.md §8d's "what eager PyTorch
can reach on the same B200".
"""
import numpy as np
import torch
import torch.nn.functional as F


class DCCFPort(torch.nn.Module):
    def __init__(self, user_num, item_num, feature_embedding, expo_prob, sample_num=10, attribute_num=2, std=0.1,
                 dim=64, seed=2019):
        super().__init__()
        torch.manual_seed(seed)
        self.user_num, self.item_num = user_num, item_num
        self.sample_num, self.attribute_num, self.std = sample_num, attribute_num, std
        self.uid_embeddings = torch.nn.Embedding(user_num, dim)
        self.iid_embeddings = torch.nn.Embedding(item_num, dim)
        self.feature_embedding = torch.as_tensor(feature_embedding, dtype=torch.float32)
        self.mlp = torch.nn.ModuleList([torch.nn.Linear(dim + self.feature_embedding.shape[1], dim)])
        self.expo_prob = torch.as_tensor(expo_prob, dtype=torch.float32)
        for m in self.modules():                                   # BaseModel.init_paras
            if type(m) == torch.nn.Linear:
                torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)
                torch.nn.init.normal_(m.bias, mean=0.0, std=0.01)
            elif type(m) == torch.nn.Embedding:
                torch.nn.init.normal_(m.weight, mean=0.0, std=0.01)

    def to_device(self, device):
        self.to(device)
        self.feature_embedding = self.feature_embedding.to(device)      # plain attributes, as in DCCF.py:55,64
        self.expo_prob = self.expo_prob.to(device)
        return self

    def predict(self, feed_dict):
        dev = self.uid_embeddings.weight.device
        u_ids = feed_dict['X'][:, 0]
        i_ids = feed_dict['X'][:, 1]
        S, A = self.sample_num, self.attribute_num
        sample_item = feed_dict.get('sample_item')
        if sample_item is None:
            sample_item = torch.randint(self.item_num, size=(u_ids.shape[0], S))                    # DCCF.py:72 (CPU)
        sample_item = sample_item.to(dev)
        items = torch.cat((i_ids.view(-1, 1), sample_item), 1)                                      # :74
        items = items.view(-1, S + 1, 1).expand(items.shape[0], S + 1, A)                           # :76
        users = u_ids.view(-1, 1, 1).expand(items.shape[0], items.shape[1], items.shape[2])         # :77
        true_items = i_ids.view(-1, 1, 1).expand(items.shape[0], items.shape[1], items.shape[2])    # :78
        uid, iid, fid = users.reshape(-1), items.reshape(-1), true_items.reshape(-1)                # :80-82
        user_embeddings = self.uid_embeddings(uid)                                                  # :84
        item_embeddings = self.iid_embeddings(iid)                                                  # :85
        feature_embeddings = self.feature_embedding[fid]                                            # :86
        noise = feed_dict.get('noise')
        if noise is None:
            noise = torch.empty(feature_embeddings.shape, device=dev).normal_(std=self.std)         # :87
        x = torch.cat((item_embeddings, feature_embeddings + noise), 1)                             # :89
        mask = feed_dict.get('dropout_mask')
        for layer in self.mlp:                                                                      # :91-94
            x = F.relu(layer(x))
            if mask is not None:
                x = x * mask
            else:
                x = torch.nn.Dropout(p=feed_dict['dropout'])(x)
        mlp_out = (user_embeddings * x).sum(dim=1).reshape(items.shape)                             # :96
        exposure_score = torch.softmax(self.expo_prob[uid, iid].reshape(items.shape), dim=1)        # :98
        prediction = (exposure_score * mlp_out).sum(dim=1).mean(dim=1).view([-1])                   # :100
        return {'prediction': prediction, 'check': [('prediction', prediction)]}

    def forward(self, feed_dict):
        out = self.predict(feed_dict)
        b = int(feed_dict['Y'].shape[0] / 2)
        pos, neg = out['prediction'][:b], out['prediction'][b:]
        out['loss'] = -(pos - neg).sigmoid().log().sum()                                            # :116-120
        return out

    def l2(self):
        return sum((p ** 2).sum() for p in self.parameters())                                       # BaseModel.py:179-187


def fit_step(model, optimizer, feed_dict, l2_weight):
    """One iteration of BaseRunner.fit (BaseRunner.py:175-188)."""
    optimizer.zero_grad()
    out = model(feed_dict)
    loss = out['loss'] + model.l2() * l2_weight
    loss.backward()
    torch.nn.utils.clip_grad_value_(model.parameters(), 50)
    optimizer.step()
    return out


def evaluate_users(pred, uid, iid, Y, k=5):
    """Host ranking metrics of BaseModel.evaluate_method (BaseModel.py:82-126) without pandas: mean ndcg@k,
    recall@k, precision@k over users."""
    from oracle import dccf_oracle as O
    _, _, _, m = O.rank_users(pred, uid, Y, iid, k)
    return float(m[:, 0].mean()), float(m[:, 3].mean()), float(m[:, 2].mean())
