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
