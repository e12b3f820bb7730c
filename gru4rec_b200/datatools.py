"""Host-side data helpers with the behaviour of the reference's datatools.py (hidasib/GRU4Rec datatools.py:12-39):
`sort_if_needed` sorts a frame in place only when it has to and prints what it did; `compute_offset` returns the
CSR offsets of the groups of a column."""
import time
import numpy as np


def _first_unsorted_column(frame, columns, any_order_first_dim):
    """Name of the first column that breaks the (session, time, ...) order, or None.

    Same test as the reference (datatools.py:15-26): the leading column must be non-decreasing -- or, with
    `any_order_first_dim`, merely grouped (as many value changes as distinct values minus one); every later column must be
    non-decreasing wherever the column just before it does not change."""
    changed_before = None
    for position, name in enumerate(columns):
        values = frame[name].to_numpy()
        changed = values[1:] != values[:-1]
        if position == 0 and any_order_first_dim:
            in_order = frame[name].nunique() == int(changed.sum()) + 1
        elif position == 0:
            in_order = bool((values[1:] >= values[:-1]).all())
        else:
            in_order = bool((changed_before | (values[1:] >= values[:-1])).all())
        if not in_order:
            return name
        changed_before = changed
    return None


def sort_if_needed(data, columns, any_order_first_dim=False):
    offender = _first_unsorted_column(data, columns, any_order_first_dim)
    if offender is None:
        print('The dataframe is already sorted by {}'.format(', '.join(columns)))
        return
    print('The dataframe is not sorted by {}, sorting now'.format(offender))
    started = time.time()
    data.sort_values(columns, inplace=True)
    print('Data is sorted in {:.2f}'.format(time.time() - started))


def compute_offset(data, column):
    """int32 offsets [0, n_0, n_0 + n_1, ...] of the groups of `column` in sorted group order (datatools.py:36-39)."""
    values = data[column].to_numpy()
    if len(values) and bool((values[1:] >= values[:-1]).all()):      # sorted (the only way fit() / evaluate_gpu call it): change points
        starts = np.flatnonzero(values[1:] != values[:-1]) + 1
        return np.concatenate([[0], starts, [len(values)]]).astype(np.int32)
    sizes = data.groupby(column).size().to_numpy()
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
