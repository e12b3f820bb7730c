"""Synthetic session data of the shapes SURVEY.md section 8(d) specifies (no datasets are available offline).

All draws come from numpy.random.RandomState(seed); the same generator feeds tests, bench.py and
oracle/make_golden.py so every arm sees identical inputs.
"""
import numpy as np
import pandas as pd


def make_sessions(n_items=1000, n_events=10000, seed=0, zipf_a=1.0, max_len=20, item_as_str=False,
                  session_key='SessionId', item_key='ItemId', time_key='Time'):
    """Sessions of length 2 + Geometric(0.5) - 1 (truncated to max_len), Zipf-like item popularity over a
    seed-42 permutation, session start times cumulative Exp(1) so sessions are time ordered."""
    rs = np.random.RandomState(seed)
    lens = []
    total = 0
    while total < n_events:
        l = min(2 + rs.geometric(0.5) - 1, max_len)
        lens.append(l)
        total += l
    lens = np.array(lens, dtype=np.int64)
    n_ev = int(lens.sum())
    p = 1.0 / (np.arange(n_items) + 1.0) ** zipf_a
    p /= p.sum()
    perm = np.random.RandomState(42).permutation(n_items)
    items = perm[rs.choice(n_items, size=n_ev, p=p)]
    sess = np.repeat(np.arange(len(lens)), lens)
    t0 = np.cumsum(rs.exponential(1.0, size=len(lens))) * 100.0
    within = np.concatenate([np.arange(l) for l in lens])
    times = np.repeat(t0, lens) + within
    ids = items + 1000
    df = pd.DataFrame({session_key: sess.astype(np.int32),
                       item_key: ids.astype(str) if item_as_str else ids.astype(np.int64),
                       time_key: times})
    return df


def train_test_split(df, test_frac=0.2, session_key='SessionId'):
    """Last `test_frac` of the sessions (they are time ordered) form the test set."""
    n = df[session_key].nunique()
    cut = int(n * (1 - test_frac))
    tr = df[df[session_key] < cut].copy()
    te = df[df[session_key] >= cut].copy()
    return tr, te


def make_session_arrays(n_items, n_events, seed=0, zipf_a=1.0, max_len=20):
    """Array form of make_sessions for large synthetic workloads (no DataFrame): returns
    (data_items int64 [events], offset_sessions int32 [sessions+1], session_order int64 [sessions], supports int64 [n_items]).
    Sessions are already (session, time) sorted and time ordered, so the order is the identity."""
    rs = np.random.RandomState(seed)
    n_sess_guess = int(n_events / 2.9) + 16
    lens = np.minimum(2 + rs.geometric(0.5, size=n_sess_guess) - 1, max_len).astype(np.int64)
    cs = np.cumsum(lens)
    n_sess = int(np.searchsorted(cs, n_events) + 1)
    lens = lens[:n_sess]
    n_ev = int(lens.sum())
    p = 1.0 / (np.arange(n_items) + 1.0) ** zipf_a
    p /= p.sum()
    perm = np.random.RandomState(42).permutation(n_items)
    items = perm[rs.choice(n_items, size=n_ev, p=p)].astype(np.int64)
    # make sure every item id occurs (the reference maps ids that occur in the data; unused ids would shrink n_items)
    items[:n_items] = np.arange(n_items)
    offset = np.zeros(n_sess + 1, dtype=np.int32)
    offset[1:] = np.cumsum(lens)
    supports = np.bincount(items, minlength=n_items).astype(np.int64)
    return items, offset, np.arange(n_sess, dtype=np.int64), supports
