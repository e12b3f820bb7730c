"""evaluate_gpu with the reference's signature and return value (hidasib/GRU4Rec evaluation.py:15-147)."""
import numpy as np
import pandas as pd

from . import _lib

_MODES = {'standard': 0, 'conservative': 1, 'median': 2, 'tiebreaking': 3}


def evaluate_gpu(gru, test_data, items=None, session_key='SessionId', item_key='ItemId', time_key='Time', cut_off=[20], batch_size=100, mode='standard'):
    '''
    Recall@N and MRR@N of next-item prediction, session-parallel (evaluation.py:15-147).
    Returns (recall_list, mrr_list), one entry per cut-off.  `mode` as in the reference; 'tiebreaking' adds U(0,1)*1e-10 to
    every score before the standard ranking (evaluation.py:55,65) -- it only matters where scores saturate at (near) zero; the
    noise is a counter hash on the device (the reference's comes from Theano's MRG stream).  `items`: the targets are ranked against these item ids only (evaluation.py:52-56,84-100); as in the
    reference the target's own score competes only if the target is listed, so 'conservative' can give rank 0 (MRR = inf).
    '''
    if gru.error_during_train: raise Exception
    if mode not in _MODES:
        raise NotImplementedError
    multi_cut_off = (type(cut_off) == list) or (type(cut_off) == tuple)
    cuts = list(cut_off) if multi_cut_off else [cut_off]
    print('Measuring Recall@{} and MRR@{}'.format(','.join([str(c) for c in cuts]), ','.join([str(c) for c in cuts])))
    test_data = pd.merge(test_data, pd.DataFrame({'ItemIdx': gru.itemidmap.values, item_key: gru.itemidmap.index}), on=item_key, how='inner')
    test_data.sort_values([session_key, time_key, item_key], inplace=True)
    test_data_items = test_data.ItemIdx.values
    offset_sessions = np.zeros(test_data[session_key].nunique() + 1, dtype=np.int32)
    offset_sessions[1:] = test_data.groupby(session_key).size().cumsum()
    eng = gru._ensure_engine(batch_size)
    sched = _lib.Schedule(test_data_items, offset_sessions, None, batch_size, 0, mode=1)
    if items is not None:
        eng.set_eval_items(gru.itemidmap[items].values)       # KeyError for unknown ids, as the reference's gru.itemidmap[items]
    try:
        rec, mrr, n = eng.eval_schedule(sched, cuts, _MODES[mode])
    finally:
        if items is not None:
            eng.set_eval_items(None)
    # the scoring hidden state is shared with predict_next_batch (gru4rec.py:696-697 keeps separate buffers): force its reset
    gru.predict = None
    recall = [float(r) / n for r in rec]
    mrrs = [float(m) / n for m in mrr]
    return recall, mrrs
