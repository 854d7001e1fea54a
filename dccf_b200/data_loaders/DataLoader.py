"""Dataset reader with the interface of the reference's src/data_loaders/DataLoader.py:13-290.

Reads `<path>/<dataset>/<dataset>.{train,validation,test}.csv` (headerless uid,iid,label,time), builds or
reads `.info.json` (column max/min -> user_num, item_num, label range) and the per-user history files
`.train_group.csv` / `.vt_group.csv`.  Runs once per job: host code, outside the kernels' scope."""
import json
import logging
import os

import numpy as np
import pandas as pd

from ..utils import global_p
from ..utils.mining import group_user_interactions_df


class DataLoader(object):
    @staticmethod
    def parse_data_args(parser):
        """Flags and defaults of DataLoader.py:19-33."""
        parser.add_argument('--path', type=str, default='../datasets/', help='Input data dir.')
        parser.add_argument('--dataset', type=str, default='ml100k-1-5', help='Choose a dataset.')
        parser.add_argument('--sep', type=str, default=',', help='sep of csv file.')
        parser.add_argument('--label', type=str, default='label', help='name of dataset label column.')
        return parser

    def __init__(self, path, dataset, label='label', load_data=True, sep='\t', seqs_sep=','):
        self.dataset = dataset
        self.path = os.path.join(path, dataset)
        base = os.path.join(self.path, dataset)
        self.train_file = base + global_p.TRAIN_SUFFIX
        self.validation_file = base + global_p.VALIDATION_SUFFIX
        self.test_file = base + global_p.TEST_SUFFIX
        self.info_file = base + global_p.INFO_SUFFIX
        self.user_file = base + global_p.USER_SUFFIX
        self.item_file = base + global_p.ITEM_SUFFIX
        self.train_his_file = base + global_p.TRAIN_GROUP_SUFFIX
        self.vt_his_file = base + global_p.VT_GROUP_SUFFIX
        self.sep, self.seqs_sep = sep, seqs_sep
        self.load_data = load_data
        self.label = label
        self.train_df = self.validation_df = self.test_df = None
        self._load_user_item()
        self._load_data()
        self._load_his()
        self._load_info()

    # ---- files -------------------------------------------------------------------------------
    def _load_user_item(self):
        """Optional tab-separated user/item feature tables (DataLoader.py:64-75); absent for DCCF."""
        self.user_df = self.item_df = None
        if self.load_data and os.path.exists(self.user_file):
            logging.info('load user csv...')
            self.user_df = pd.read_csv(self.user_file, sep='\t')
        if self.load_data and os.path.exists(self.item_file):
            logging.info('load item csv...')
            self.item_df = pd.read_csv(self.item_file, sep='\t')

    def _load_data(self):
        """The three interaction files (DataLoader.py:77-98)."""
        cols = [global_p.UID, global_p.IID, global_p.LABEL, global_p.TIME]
        for attr, path, name in (('train_df', self.train_file, 'train'),
                                 ('validation_df', self.validation_file, 'validation'),
                                 ('test_df', self.test_file, 'test')):
            if self.load_data and os.path.exists(path):
                logging.info('load %s csv...' % name)
                df = pd.read_csv(path, sep=self.sep, names=cols)
                setattr(self, attr, df)
                logging.info('size of %s: %d' % (name, len(df)))

    def _load_info(self):
        """column_max / column_min, user_num = max uid + 1, item_num = max iid + 1
        (DataLoader.py:100-160); the json file is written on first use and read afterwards."""
        if not os.path.exists(self.info_file):
            max_dict, min_dict = {}, {}
            for df in (self.train_df, self.validation_df, self.test_df, self.user_df, self.item_df):
                if df is None:
                    continue
                for c in df.columns:
                    hi, lo = df[c].max(), df[c].min()
                    max_dict[c] = hi if c not in max_dict else max(hi, max_dict[c])
                    min_dict[c] = lo if c not in min_dict else min(lo, min_dict[c])

            def as_json(o):
                if isinstance(o, np.integer):
                    return int(o)
                raise TypeError

            with open(self.info_file, 'w') as f:
                f.write(json.dumps(max_dict, default=as_json) + os.linesep + json.dumps(min_dict, default=as_json))
        else:
            with open(self.info_file) as f:
                lines = f.readlines()
            max_dict, min_dict = json.loads(lines[0]), json.loads(lines[1])
        self.column_max, self.column_min = max_dict, min_dict
        self.label_max, self.label_min = self.column_max[self.label], self.column_min[self.label]
        logging.info('label: %d-%d' % (self.label_min, self.label_max))
        self.user_num = self.column_max['uid'] + 1 if 'uid' in self.column_max else 0
        self.item_num = self.column_max['iid'] + 1 if 'iid' in self.column_max else 0
        logging.info('# of users: %d' % self.user_num)
        logging.info('# of items: %d' % self.item_num)
        for kind, prefix in (('user', 'u_'), ('item', 'i_'), ('context', 'c_')):
            found = [f for f in self.column_max if f.startswith(prefix)]
            setattr(self, kind + '_features', found)
            logging.info('# of %s features: %d' % (kind, len(found)))
        self.features = self.context_features + self.user_features + self.item_features
        logging.info('# of features: %d' % len(self.features))

    def _load_his(self):
        """Per-user positive-item lists for train and validation+test (DataLoader.py:162-194)."""
        self.train_his_df = self.train_user_his = None
        self.vt_his_df = self.vt_user_his = None
        if not self.load_data:
            return
        if not os.path.exists(self.train_his_file):
            logging.info('building train history csv...')
            group_user_interactions_df(self.train_df, label=self.label, seq_sep=self.seqs_sep) \
                .to_csv(self.train_his_file, index=False, sep=self.sep)
        if not os.path.exists(self.vt_his_file):
            logging.info('building vt history csv...')
            vt_df = pd.concat([self.validation_df, self.test_df])
            group_user_interactions_df(vt_df, label=self.label, seq_sep=self.seqs_sep) \
                .to_csv(self.vt_his_file, index=False, sep=self.sep)

        def to_dict(his_df):
            return {int(u): [int(j) for j in str(s).split(self.seqs_sep)]
                    for u, s in zip(his_df['uid'].tolist(), his_df['iids'].tolist())}

        logging.info('load history csv...')
        self.train_his_df = pd.read_csv(self.train_his_file, sep=self.sep)
        self.train_user_his = to_dict(self.train_his_df)
        self.vt_his_df = pd.read_csv(self.vt_his_file, sep=self.sep)
        self.vt_user_his = to_dict(self.vt_his_df)

    # ---- model-facing ------------------------------------------------------------------------
    def feature_info(self, include_id=True, include_item_features=True, include_user_features=True):
        """Feature names, total multi-hot width and per-feature index ranges (DataLoader.py:196-224)."""
        features = []
        if include_id:
            features += ['uid', 'iid']
        if include_user_features:
            features += self.user_features
        if include_item_features:
            features += self.item_features
        dims, lo, hi = 0, [], []
        for f in features:
            lo.append(dims)
            dims += int(self.column_max[f] + 1)
            hi.append(dims - 1)
        logging.info('Model # of features %d' % len(features))
        logging.info('Model # of feature dims %d' % dims)
        return features, dims, lo, hi

    def drop_neg(self):
        """Top-n recommendation keeps label > 0 only, relabelled 1 (DataLoader.py:276-290)."""
        logging.info('Drop Neg Samples...')
        for attr in ('train_df', 'validation_df', 'test_df'):
            df = getattr(self, attr)
            df = df[df[self.label] > 0].reset_index(drop=True)
            df[self.label] = 1
            setattr(self, attr, df)
        logging.info('size of train: %d' % len(self.train_df))
        logging.info('size of validation: %d' % len(self.validation_df))
        logging.info('size of test: %d' % len(self.test_df))
