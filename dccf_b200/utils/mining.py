"""`group_user_interactions_df` of src/utils/mining.py:18-29 (the only function of that module on the
DCCF path; the association-rule miner needs `pymining` and is out of scope)."""
import pandas as pd


def group_user_interactions_df(in_df, label='label', seq_sep=','):
    df = in_df
    if label in df.columns:
        df = df[df[label] > 0]
    uids, inters = [], []
    for uid, group in df.groupby('uid'):
        uids.append(uid)
        inters.append(seq_sep.join(str(i) for i in group['iid'].tolist()))
    out = pd.DataFrame()
    out['uid'] = uids
    out['iids'] = inters
    return out
