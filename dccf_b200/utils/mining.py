"""`group_user_interactions_df` of src/utils/mining.py:18-29 (the only function of that module on the
DCCF path; the association-rule miner needs `pymining` and is out of scope)."""
import numpy as np
import pandas as pd


def group_user_interactions_df(in_df, label='label', seq_sep=','):
    """One row per user (ascending uid): the user's positively labelled items in file order, joined by seq_sep.
    Same frame as the reference's `for uid, group in df.groupby('uid')` loop; grouped here by one stable sort
    instead of one DataFrame per user (seconds at 48 k users)."""
    df = in_df
    if label in df.columns:
        df = df[df[label] > 0]
    uid = df['uid'].to_numpy()
    iid = df['iid'].to_numpy()
    order = np.argsort(uid, kind='stable')
    users, starts = np.unique(uid[order], return_index=True)
    ends = np.append(starts[1:], len(order))
    items = [str(i) for i in iid[order].tolist()]
    out = pd.DataFrame()
    out['uid'] = users
    out['iids'] = [seq_sep.join(items[a:b]) for a, b in zip(starts.tolist(), ends.tolist())]
    return out
