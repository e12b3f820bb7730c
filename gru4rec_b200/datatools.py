"""Host-side data helpers with the reference's behaviour (datatools.py:12-39 of hidasib/GRU4Rec)."""
import time
import numpy as np


def sort_if_needed(data, columns, any_order_first_dim=False):
    """In-place sort of `data` by `columns` unless it is already sorted (datatools.py:12-34); same prints."""
    is_sorted = True
    neq_masks = []
    col = columns[0]
    for i, col in enumerate(columns):
        vals = data[col].values
        neq_masks.append(vals[1:] != vals[:-1])
        if i == 0:
            if any_order_first_dim:
                is_sorted = is_sorted and (data[col].nunique() == neq_masks[0].sum() + 1)
            else:
                is_sorted = is_sorted and bool(np.all(vals[1:] >= vals[:-1]))
        else:
            is_sorted = is_sorted and bool(np.all(neq_masks[i - 1] | (vals[1:] >= vals[:-1])))
        if not is_sorted:
            break
    if is_sorted:
        print('The dataframe is already sorted by {}'.format(', '.join(columns)))
    else:
        print('The dataframe is not sorted by {}, sorting now'.format(col))
        t0 = time.time()
        data.sort_values(columns, inplace=True)
        t1 = time.time()
        print('Data is sorted in {:.2f}'.format(t1 - t0))


def compute_offset(data, column):
    """CSR-style session offsets, int32 (datatools.py:36-39)."""
    offset = np.zeros(data[column].nunique() + 1, dtype=np.int32)
    offset[1:] = data.groupby(column).size().cumsum()
    return offset
