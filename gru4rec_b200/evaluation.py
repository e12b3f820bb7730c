"""evaluate_gpu with the reference's signature and return value (hidasib/GRU4Rec evaluation.py:15-147)."""
import os

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
    Under a torch.distributed job (torchrun, one process per GPU) every rank calls this with the same test data and scores
    every world-th session; all ranks return the same (summed) result.  G4R_EVAL_SHARD=0 switches the sharding off.
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
    n_sessions = len(offset_sessions) - 1
    world, rank = gru._world()
    if world > 1 and os.environ.get('G4R_EVAL_SHARD', '1') == '0':
        world, rank = 1, 0                                     # every rank scores the whole test set on its own replica
    if world > 1 and n_sessions < batch_size:
        # the reference indexes offset_sessions[arange(batch_size) + 1] (evaluation.py:93-94): same error, whatever the shard sizes
        raise IndexError('index out of bounds: fewer sessions than batch_size (reference: IndexError at evaluation.py:94)')
    eng = gru._ensure_engine(batch_size)
    if items is not None:
        eng.set_eval_items(gru.itemidmap[items].values)       # KeyError for unknown ids, as the reference's gru.itemidmap[items]
    try:
        if world == 1:
            sched = _lib.Schedule(test_data_items, offset_sessions, None, batch_size, 0, mode=1)
            rec, mrr, n = eng.eval_schedule(sched, cuts, _MODES[mode])
        else:
            # one process per GPU: rank r scores every world-th session on its full replica of the model; the hit and
            # reciprocal-rank sums (double) and the event count are summed over the ranks -- no exchange on the data path
            from .parallel import shard_eval_sessions, allreduce_sum
            import torch.distributed as dist
            mine = shard_eval_sessions(n_sessions, rank, world)
            rec, mrr, n = np.zeros(len(cuts)), np.zeros(len(cuts)), 0
            if len(mine):
                sched = _lib.Schedule(test_data_items, offset_sessions, mine, min(batch_size, len(mine)), 0, mode=1)
                rec, mrr, n = eng.eval_schedule(sched, cuts, _MODES[mode])
            tot = allreduce_sum(np.concatenate([rec, mrr, [float(n)]]), dist)
            rec, mrr, n = tot[:len(cuts)], tot[len(cuts):2 * len(cuts)], int(round(tot[-1]))
    finally:
        if items is not None:
            eng.set_eval_items(None)
    # the scoring hidden state is shared with predict_next_batch (gru4rec.py:696-697 keeps separate buffers): force its reset
    gru.predict = None
    recall = [float(r) / n for r in rec]
    mrrs = [float(m) / n for m in mrr]
    return recall, mrrs
