"""Batch producer / negative sampler with the interface and the exact sampling semantics of the
reference's src/data_processor/DataProcessor.py:15-524.

Index parity contract: every `np.random` call the reference makes is made here in
the same order with the same arguments — `np.random.randint(item_num)` rejection draws
(DataProcessor.py:498-504), the `np.random.choice(remain, neg_n, replace=False)` branch when fewer than
20 % of the items remain (DP:490-493,505-512) and the replayed-state epoch shuffle
(src/utils/utils.py:82-92) — so negatives and batch contents are bit-identical under the same seed
(checked against fixtures produced by the reference: tests/golden/sampler.npz).

What is different is the data movement: the reference uploads every tensor of every batch separately
(`numpy_to_torch` per field, DP:138-207); here one epoch is assembled into a single [rows,2] int64 array,
uploaded once, and the per-batch feed dicts hold views of it.
"""
import logging
from collections import defaultdict

import numpy as np
import pandas as pd
import torch

from ..utils import global_p, utils


class DataProcessor(object):
    data_columns = ['X']

    @staticmethod
    def parse_dp_args(parser):
        """--test_neg_n, default 100 (DataProcessor.py:19-27)."""
        parser.add_argument('--test_neg_n', type=int, default=100,
                            help='Negative sample num for each instance in test/validation set.')
        return parser

    def __init__(self, data_loader, model, rank, test_neg_n):
        self.data_loader = data_loader
        self.model = model
        self.rank = rank
        self.test_neg_n = test_neg_n
        self.train_data = self.validation_data = self.test_data = None
        if self.rank == 1:
            # per-user positive sets: what a negative must avoid (DataProcessor.py:44-53)
            self.train_history_dict = defaultdict(set)
            for uid, items in data_loader.train_user_his.items():
                self.train_history_dict[uid] = set(items)
            self.vt_history_dict = defaultdict(set)
            for uid, items in data_loader.vt_user_his.items():
                self.vt_history_dict[uid] = set(items)
        self.vt_batches_buffer = {}
        self.use_native_sampler = True      # dccf_sample_negatives (C++, draw-for-draw identical); False = Python loop
        self.fast_ids_path = True           # id-only models: data dicts / epochs from arrays; False = DataFrame route
        self._csr = None

    # ---- data dicts --------------------------------------------------------------------------
    def get_train_data(self, epoch):
        """Train dict; shuffled in place (cumulatively) when epoch >= 0 (DataProcessor.py:57-71)."""
        if self.train_data is None or epoch < 0:
            logging.info('Prepare Train Data...')
            self.train_data = self.format_data_dict(self.data_loader.train_df)
            self.train_data[global_p.K_SAMPLE_ID] = np.arange(0, len(self.train_data['Y']))
        if epoch >= 0:
            utils.shuffle_in_unison_scary(self.train_data)
        return self.train_data

    def _ids_only(self, df):
        """True when format_data_dict's 'X' is exactly [uid, iid] for this model (DCCF: append_id, no id / user / item /
        context feature columns, RecModel.py:10-13): the data dicts can then be assembled from arrays directly — same
        contents as the DataFrame route (generate_neg_df + concat + format_data_dict), checked by
        tests/test_host_parity.py — without a 48 M-row pandas merge for 1000 negatives x 48 k users."""
        m, l = self.model, self.data_loader
        return (m.append_id and not m.include_id and not m.include_context_features
                and not (l.user_df is not None and m.include_user_features)
                and not (l.item_df is not None and m.include_item_features)
                and 'uid' in df and 'iid' in df and l.label in df.columns)

    def _negatives(self, uids, neg_n, train):
        """int64 [len(uids) * neg_n] negatives, user-major — the arrays behind _sample_neg_from_uid_list's frame."""
        if self.use_native_sampler and len(uids) > 0:
            return self._sample_neg_native(np.asarray(uids, dtype=np.int64), neg_n, train)
        return self._sample_neg_from_uid_list(uids=uids, neg_n=neg_n, train=train)['iid_neg'].to_numpy(dtype=np.int64)

    def _vt_data(self, df):
        if self.rank == 1 and self.fast_ids_path and self._ids_only(df):
            uid = df['uid'].to_numpy(dtype=np.int64)
            iid = df['iid'].to_numpy(dtype=np.int64)
            _, first = np.unique(uid, return_index=True)
            f_u = uid[np.sort(first)]                    # distinct users in first-appearance order (DP:420-430)
            neg = self._negatives(f_u, self.test_neg_n, train=False)
            data = {'uid': np.concatenate([uid, np.repeat(f_u, self.test_neg_n)]), 'iid': np.concatenate([iid, neg])}
            data['Y'] = np.concatenate([np.asarray(df[self.data_loader.label], dtype=np.float32),
                                        np.zeros(len(neg), dtype=np.float32)])
            data['X'] = np.stack([data['uid'], data['iid']], axis=1)
            data[global_p.K_SAMPLE_ID] = np.arange(0, len(data['Y']))
            return data
        if self.rank == 1:
            neg_df = self.generate_neg_df(uid_list=df['uid'].tolist(), iid_list=df['iid'].tolist(), df=df,
                                          neg_n=self.test_neg_n, train=False)
            df = pd.concat([df, neg_df], ignore_index=True)
        data = self.format_data_dict(df)
        data[global_p.K_SAMPLE_ID] = np.arange(0, len(data['Y']))
        return data

    def get_validation_data(self):
        """Validation positives + test_neg_n negatives per distinct user, built once (DP:73-90)."""
        if self.validation_data is None:
            logging.info('Prepare Validation Data...')
            self.validation_data = self._vt_data(self.data_loader.validation_df)
        return self.validation_data

    def get_test_data(self):
        """Test positives + test_neg_n negatives per distinct user, built once (DP:92-111)."""
        if self.test_data is None:
            logging.info('Prepare Test Data...')
            self.test_data = self._vt_data(self.data_loader.test_df)
        return self.test_data

    def get_train_batches(self, batch_size, epoch):
        return self.prepare_batches(self.get_train_data(epoch), batch_size, train=True)

    def get_validation_batches(self, batch_size):
        return self.prepare_batches(self.get_validation_data(), batch_size, train=False)

    def get_test_batches(self, batch_size):
        return self.prepare_batches(self.get_test_data(), batch_size, train=False)

    # ---- feed dicts --------------------------------------------------------------------------
    def _get_feed_dict_rt(self, data, batch_start, batch_size, train):
        """Plain slice of the data dict (DataProcessor.py:138-158)."""
        end = min(len(data['X']), batch_start + batch_size)
        real = end - batch_start
        feed = {'train': train, 'rank': 0, global_p.K_SAMPLE_ID: data[global_p.K_SAMPLE_ID][batch_start:end]}
        y = data['Y'][batch_start:end] if 'Y' in data else np.zeros(shape=real)
        feed['Y'] = utils.numpy_to_torch(y)
        for c in self.data_columns:
            feed[c] = utils.numpy_to_torch(data[c][batch_start:end])
        return feed

    def _get_feed_dict_rk(self, data, batch_start, batch_size, train, neg_data=None):
        """Top-n batch: eval = plain slice with rank 1; train = [positives ; their negatives]
        (DataProcessor.py:160-207)."""
        if not train:
            feed = self._get_feed_dict_rt(data, batch_start, batch_size, train)
            feed['rank'] = 1
            return feed
        end = min(len(data['X']), batch_start + batch_size)
        real = end - batch_start
        if neg_data is None:
            logging.warning('neg_data is None')
            neg_df = self.generate_neg_df(uid_list=data['uid'][batch_start:end], iid_list=data['iid'][batch_start:end],
                                          df=self.data_loader.train_df, neg_n=1, train=True)
            neg_cols = {c: self.format_data_dict(neg_df)[c] for c in self.data_columns}
        else:
            neg_cols = {c: neg_data[c][batch_start:end] for c in self.data_columns}
        y = np.concatenate([np.ones(real, dtype=np.float32), np.zeros(real, dtype=np.float32)])
        sample_id = data[global_p.K_SAMPLE_ID][batch_start:end]
        feed = {'train': train, 'rank': 1, 'Y': utils.numpy_to_torch(y),
                global_p.K_SAMPLE_ID: np.concatenate([sample_id, sample_id + len(self.train_data['Y'])]),
                global_p.REAL_BATCH_SIZE: real, global_p.TOTAL_BATCH_SIZE: real * 2}
        for c in self.data_columns:
            feed[c] = utils.numpy_to_torch(np.concatenate([data[c][batch_start:end], neg_cols[c]]))
        return feed

    def get_feed_dict(self, data, batch_start, batch_size, train, neg_data=None):
        if self.rank == 1:
            return self._get_feed_dict_rk(data, batch_start, batch_size, train, neg_data)
        return self._get_feed_dict_rt(data, batch_start, batch_size, train)

    # ---- whole-epoch batch lists -----------------------------------------------------------------
    def _epoch_views(self, rows_X, rows_Y, bounds):
        """One host->device copy for the whole list, feed dicts get views (replaces the per-field
        `numpy_to_torch` uploads of DP:138-207)."""
        X = torch.from_numpy(np.ascontiguousarray(rows_X))
        Y = torch.from_numpy(np.ascontiguousarray(rows_Y))
        if torch.cuda.device_count() > 0:
            X = X.pin_memory().cuda(non_blocking=True)
            Y = Y.pin_memory().cuda(non_blocking=True)
        return [(X[a:b], Y[a:b]) for a, b in bounds]

    def _prepare_batches_rt(self, data, batch_size, train):
        """Rating/click prediction batches (DataProcessor.py:209-225)."""
        if data is None:
            return None
        n = len(data['X'])
        assert n > 0
        total = (n + batch_size - 1) // batch_size
        bounds = [(k * batch_size, min(n, (k + 1) * batch_size)) for k in range(total)]
        y = data['Y'] if 'Y' in data else np.zeros(n, dtype=np.float32)
        views = self._epoch_views(np.asarray(data['X']), np.asarray(y), bounds)
        batches = []
        for (a, b), (xv, yv) in zip(bounds, views):
            batches.append({'train': train, 'rank': 0, global_p.K_SAMPLE_ID: data[global_p.K_SAMPLE_ID][a:b],
                            'Y': yv, 'X': xv})
        return batches

    def _prepare_batches_rk(self, data, batch_size, train):
        """Top-n batches; for training one negative per positive is drawn for the WHOLE epoch first, in
        the current row order (DataProcessor.py:227-250)."""
        if data is None:
            return None
        n = len(data['X'])
        assert n > 0
        if not train:
            batches = self._prepare_batches_rt(data, batch_size, train)
            for b in batches:
                b['rank'] = 1
            return batches
        pos_X = np.asarray(data['X'])
        if self.fast_ids_path and self._ids_only(self.data_loader.train_df) and pos_X.shape[1] == 2:
            neg_X = np.stack([data['uid'], self._negatives(data['uid'], 1, train=True)], axis=1).astype(pos_X.dtype)
        else:
            neg_df = self.generate_neg_df(uid_list=data['uid'], iid_list=data['iid'], df=self.data_loader.train_df,
                                          neg_n=1, train=True)
            neg_X = self.format_data_dict(neg_df)['X']
        # epoch layout: batch k = [its positives ; their negatives] (DP:160-207), all batches back to back
        total = (n + batch_size - 1) // batch_size
        n_full = n // batch_size
        w = pos_X.shape[1]
        n_train = len(self.train_data['Y'])
        sid = np.asarray(data[global_p.K_SAMPLE_ID])
        rows_X = np.empty((2 * n, w), dtype=pos_X.dtype)
        rows_Y = np.empty(2 * n, dtype=np.float32)
        rows_sid = np.empty(2 * n, dtype=sid.dtype)
        m = n_full * batch_size
        if n_full:
            vx = rows_X[:2 * m].reshape(n_full, 2, batch_size, w)
            vx[:, 0] = pos_X[:m].reshape(n_full, batch_size, w)
            vx[:, 1] = neg_X[:m].reshape(n_full, batch_size, w)
            vy = rows_Y[:2 * m].reshape(n_full, 2, batch_size)
            vy[:, 0] = 1.0
            vy[:, 1] = 0.0
            vs = rows_sid[:2 * m].reshape(n_full, 2, batch_size)
            vs[:, 0] = sid[:m].reshape(n_full, batch_size)
            vs[:, 1] = vs[:, 0] + n_train
        if m < n:                                       # ragged last batch
            real = n - m
            rows_X[2 * m:2 * m + real] = pos_X[m:]
            rows_X[2 * m + real:] = neg_X[m:]
            rows_Y[2 * m:2 * m + real] = 1.0
            rows_Y[2 * m + real:] = 0.0
            rows_sid[2 * m:2 * m + real] = sid[m:]
            rows_sid[2 * m + real:] = sid[m:] + n_train
        spans = [(k * batch_size, min(n, (k + 1) * batch_size)) for k in range(total)]
        bounds = [(2 * a, 2 * b) for a, b in spans]
        views = self._epoch_views(rows_X, rows_Y, bounds)
        return [{'train': train, 'rank': 1, 'Y': yv, 'X': xv, global_p.K_SAMPLE_ID: rows_sid[2 * a:2 * b],
                 global_p.REAL_BATCH_SIZE: b - a, global_p.TOTAL_BATCH_SIZE: 2 * (b - a)}
                for (a, b), (xv, yv) in zip(spans, views)]

    def prepare_batches(self, data, batch_size, train):
        """All batches of a data dict; validation/test lists are cached (DataProcessor.py:252-275)."""
        key = ''
        if data is self.validation_data:
            key = 'validation_' + str(batch_size)
        elif data is self.test_data:
            key = 'test_' + str(batch_size)
        if key in self.vt_batches_buffer:
            return self.vt_batches_buffer[key]
        if self.rank == 1:
            batches = self._prepare_batches_rk(data, batch_size, train)
        else:
            batches = self._prepare_batches_rt(data, batch_size, train)
        if key:
            self.vt_batches_buffer[key] = batches
        return batches

    # ---- DataFrame -> data dict ------------------------------------------------------------------
    def format_data_dict(self, df):
        """uid / iid / Y / X of a DataFrame; X = [uid, iid] (+ offset feature ids when the model uses
        side features) as int64 (DataProcessor.py:292-356)."""
        loader, model = self.data_loader, self.model
        data, id_cols = {}, []
        for c in ('uid', 'iid'):
            if c in df:
                id_cols.append(c)
                data[c] = np.array(df[c].values, copy=True)
        if loader.label in df.columns:
            data['Y'] = np.array(df[loader.label], dtype=np.float32)
        else:
            logging.warning('No Labels In Data: ' + loader.label)
            data['Y'] = np.zeros(len(df), dtype=np.float32)
        ui = df[id_cols]
        out = ui
        if loader.user_df is not None and model.include_user_features:
            out = pd.merge(out, loader.user_df, on='uid', how='left')
        if loader.item_df is not None and model.include_item_features:
            out = pd.merge(out, loader.item_df, on='iid', how='left')
        out = out.fillna(0)
        if model.include_context_features:
            out = pd.concat([out, df[loader.context_features]], axis=1, ignore_index=True)
        if not model.include_id:
            out = out.drop(columns=['uid', 'iid'])
        base = 0
        shifted = {}
        for f in out.columns:
            shifted[f] = out[f].values + base
            base += int(loader.column_max[f] + 1)
        feats = np.stack([shifted[f] for f in out.columns], axis=1) if len(out.columns) else \
            np.zeros((len(df), 0), dtype=np.int64)
        if model.append_id:
            data['X'] = np.concatenate([ui.values, feats], axis=1).astype(int)
        else:
            data['X'] = feats.astype(int)
        assert len(data['X']) == len(data['Y'])
        return data

    # ---- negative sampling -----------------------------------------------------------------------
    def generate_neg_df(self, uid_list, iid_list, df, neg_n, train):
        """DataFrame of sampled negatives with df's columns and label 0 (DataProcessor.py:408-444).
        Evaluation: neg_n per DISTINCT user, in first-appearance order (DP:420-430)."""
        if not train:
            seen, f_u, f_i = set(), [], []
            for u, i in zip(uid_list, iid_list):
                if u not in seen:
                    seen.add(u)
                    f_u.append(u)
                    f_i.append(i)
        else:
            f_u, f_i = uid_list, iid_list
        neg_df = self._sample_neg_from_uid_list(uids=f_u, neg_n=neg_n, train=train, other_infos={'iid': f_i})
        other_cols = [c for c in df.columns if c not in ('uid', 'iid')]
        if other_cols:
            # the positive row's remaining columns travel with its negatives (DP:437-442)
            first = df.drop_duplicates(subset=['uid', 'iid'])
            neg_df = pd.merge(neg_df, first, on=['uid', 'iid'], how='left')
        neg_df = neg_df.drop(columns=['iid']).rename(columns={'iid_neg': 'iid'})
        neg_df = neg_df[list(df.columns)]
        neg_df[self.data_loader.label] = 0
        return neg_df

    def _history_csr(self):
        """Per-user sorted item lists of the train and validation/test histories as CSR arrays."""
        if self._csr is None:
            n_users = int(self.data_loader.user_num)

            def build(hist):
                off = np.zeros(n_users + 1, dtype=np.int64)
                for u, items in hist.items():
                    if 0 <= u < n_users:
                        off[u + 1] = len(items)
                off = np.cumsum(off)
                flat = np.zeros(max(int(off[-1]), 1), dtype=np.int64)
                for u, items in hist.items():
                    if 0 <= u < n_users and items:
                        flat[off[u]:off[u + 1]] = sorted(items)
                return off, flat

            self._csr = (n_users,) + build(self.train_history_dict) + build(self.vt_history_dict)
        return self._csr

    def _sample_neg_native(self, uids, neg_n, train):
        """dccf_sample_negatives: the same draws from numpy's global MT19937 state, in C++."""
        import ctypes
        from .. import _lib
        lib = _lib.load()
        n_users, t_off, t_items, v_off, v_items = self._history_csr()
        uids = np.ascontiguousarray(uids, dtype=np.int64)
        kind, key, pos, has_gauss, cached = np.random.get_state()
        assert kind == 'MT19937'
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        pos_c = ctypes.c_int32(int(pos))
        out = np.empty(len(uids) * neg_n, dtype=np.int64)

        def p(a):
            return a.ctypes.data_as(ctypes.c_void_p)

        rc = lib.dccf_sample_negatives(p(key), ctypes.byref(pos_c), p(uids), len(uids), int(neg_n), int(bool(train)),
                                       int(self.data_loader.item_num), n_users, p(t_off), p(t_items), p(v_off),
                                       p(v_items), p(out))
        np.random.set_state((kind, key, int(pos_c.value), has_gauss, cached))
        if rc != 0:
            raise AssertionError(lib.dccf_last_error().decode())
        return out

    def _sample_neg_from_uid_list(self, uids, neg_n, train, other_infos=None):
        """The rejection sampler (DataProcessor.py:446-524), draw for draw."""
        other_infos = other_infos or {}
        if self.use_native_sampler and len(uids) > 0:
            uid_arr = np.asarray(uids, dtype=np.int64)
            neg_df = pd.DataFrame({'uid': np.repeat(uid_arr, neg_n),
                                   'iid_neg': self._sample_neg_native(uid_arr, neg_n, train)})
            for info, values in other_infos.items():
                neg_df[info] = np.repeat(np.asarray(values), neg_n)
            return neg_df
        item_num = self.data_loader.item_num
        randint, choice = np.random.randint, np.random.choice
        train_hist, vt_hist = self.train_history_dict, self.vt_history_dict
        out_u, out_i = [], []
        drawn = defaultdict(set)            # negatives already handed to a user (kept across the epoch when train)
        for uid in uids:
            if train:
                taken = train_hist[uid] | drawn[uid]
            else:
                taken = train_hist[uid] | vt_hist[uid] | drawn[uid]
            remain = item_num - len(taken)
            pool = None
            if 1.0 * remain / item_num < 0.2:
                pool = [i for i in range(1, item_num) if i not in taken]      # item 0 excluded here only (DP:493)
            assert remain >= neg_n
            mine = drawn[uid]
            if pool is None:
                for _ in range(neg_n):
                    iid = randint(item_num)
                    while iid in taken or iid in mine:
                        iid = randint(item_num)
                    out_u.append(uid)
                    out_i.append(iid)
                    mine.add(iid)
            else:
                iids = choice(pool, neg_n, replace=False)
                out_u.extend([uid] * neg_n)
                out_i.extend(iids)
                mine.update(iids)
            if not train:
                drawn = defaultdict(set)
        neg_df = pd.DataFrame({'uid': np.asarray(out_u, dtype=np.int64), 'iid_neg': np.asarray(out_i, dtype=np.int64)})
        for info, values in other_infos.items():
            neg_df[info] = np.repeat(np.asarray(values), neg_n)
        return neg_df
