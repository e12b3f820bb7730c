"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/gru4rec.py, evaluation.py, gpu_ops.py, datatools.py) on top of oracle/theano_shim.

Run here (the GPU box has no /root/reference):   python oracle/make_golden.py
Each fixture holds, for one small configuration: the synthetic data, the reference's initial and final
weights, every train_function call (X, Y, M, R, cost), the sample store contents, dropout masks, the
epoch losses the reference printed, and evaluate_gpu's Recall/MRR.  tests/test_oracle_golden.py
replays them through oracle/gru4rec_oracle.py; tests/test_gpu_golden.py through the CUDA path.
"""
import io
import os
import sys
import re
import contextlib
import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
REF = '/root/reference'

import theano_shim
theano_shim.install()
sys.path.insert(0, REF)
cwd = os.getcwd()
import gru4rec as ref_gru4rec          # the reference module (it chdir()s during import and back)
import evaluation as ref_evaluation
os.chdir(cwd)
from gru4rec_b200.synth import make_sessions, train_test_split

CONFIGS = {
    # name: (data kwargs, model kwargs, fit kwargs)
    # BASELINE.json configs[1] at its real shape: B = 32, GRU(100), BPR-max, 2048 negative samples (param_samples/rsc15_bpr-max.py), 66 mini-batches
    'bprmax_headline_shape': (dict(n_items=1500, n_events=3600, seed=21),
                              dict(loss='bpr-max', final_act='elu-0.5', hidden_act='tanh', layers=[100], batch_size=32, n_epochs=1,
                                   learning_rate=0.05, momentum=0.3, n_sample=2048, sample_alpha=0.0, bpreg=1.0),     # lr: stable with ~1.4 duplicates per item and step
                              dict(sample_store=2048 * 80)),
    # the shape family of configs[2] (paramfiles/rees46_xe_shared_best.py: shared embedding, XE + logQ, momentum 0), dropout off for the CUDA replay
    'xe_shared_logq_l64_b48': (dict(n_items=300, n_events=2400, seed=22),
                               dict(loss='cross-entropy', final_act='softmax', layers=[64], batch_size=48, n_epochs=1, constrained_embedding=True,
                                    learning_rate=0.065, momentum=0.0, n_sample=256, sample_alpha=0.5, bpreg=0.0, logq=1.0),
                               dict(sample_store=256 * 40)),
    # the shape family of configs[3] (paramfiles/retailrocket_bprmax_shared_best.py with three layers), dropout off for the CUDA replay
    'bprmax_shared_3x32_b16': (dict(n_items=200, n_events=1500, seed=23),
                               dict(loss='bpr-max', final_act='elu-0.5', layers=[32, 32, 32], batch_size=16, n_epochs=1, constrained_embedding=True,
                                    learning_rate=0.05, momentum=0.4, n_sample=128, sample_alpha=0.4, bpreg=1.95),
                               dict(sample_store=128 * 70)),
    'bprmax_none': (dict(n_items=60, n_events=700, seed=1),
                    dict(loss='bpr-max', final_act='elu-0.5', hidden_act='tanh', layers=[12], batch_size=6, n_epochs=2,
                         learning_rate=0.2, momentum=0.3, n_sample=16, sample_alpha=0.0, bpreg=1.0),
                    dict(sample_store=16 * 40)),
    'xe_shared_logq': (dict(n_items=80, n_events=800, seed=2),
                       dict(loss='cross-entropy', final_act='softmax', layers=[12], batch_size=5, n_epochs=2,
                            constrained_embedding=True, learning_rate=0.2, momentum=0.2, n_sample=24, sample_alpha=0.5,
                            bpreg=0.0, logq=1.0, dropout_p_hidden=0.4),
                       dict(sample_store=24 * 50)),
    'xe_embed_2layer': (dict(n_items=70, n_events=700, seed=3),
                        dict(loss='cross-entropy', final_act='softmax', layers=[8, 12], batch_size=4, n_epochs=2,
                             embedding=8, learning_rate=0.1, momentum=0.0, n_sample=12, sample_alpha=0.75,
                             dropout_p_embed=0.3, dropout_p_hidden=0.2, lmbd=0.001),
                        dict(sample_store=12 * 30)),
    'top1max_none_2layer': (dict(n_items=50, n_events=600, seed=4),
                            dict(loss='top1-max', final_act='tanh', hidden_act='tanh', layers=[8, 8], batch_size=5, n_epochs=1,
                                 learning_rate=0.1, momentum=0.1, n_sample=10, sample_alpha=1.0),
                            dict(sample_store=10 * 25)),
    'bprmax_shared_nosample': (dict(n_items=40, n_events=500, seed=5),
                               dict(loss='bpr-max', final_act='elu-1', layers=[10], batch_size=6, n_epochs=1,
                                    constrained_embedding=True, learning_rate=0.05, momentum=0.4, n_sample=0, bpreg=1.95),
                               dict(sample_store=0)),
    'xe_none_default': (dict(n_items=100, n_events=900, seed=6),        # BASELINE configs[0] shape, shrunk
                        dict(loss='cross-entropy', final_act='softmax', layers=[16], batch_size=8, n_epochs=1, n_sample=32),
                        dict(sample_store=32 * 20)),
    # ---- the rest of the option surface ----
    'bpr_none_relu': (dict(n_items=50, n_events=600, seed=7),
                      dict(loss='bpr', final_act='linear', hidden_act='relu', layers=[10], batch_size=5, n_epochs=1,
                           learning_rate=0.05, momentum=0.0, n_sample=12, sample_alpha=0.5),
                      dict(sample_store=12 * 30)),
    'top1_embed_selu': (dict(n_items=55, n_events=650, seed=8),
                        dict(loss='top1', final_act='selu-1.05-1.67', hidden_act='elu-1.0', layers=[9], batch_size=6, n_epochs=1,
                             embedding=6, learning_rate=0.05, momentum=0.2, n_sample=10, sample_alpha=0.75, lmbd=0.0005),
                        dict(sample_store=10 * 30)),
    'xelogit_shared_leaky': (dict(n_items=45, n_events=600, seed=9),
                             dict(loss='xe_logit', final_act='softmax_logit', hidden_act='leaky-0.2', layers=[10], batch_size=5, n_epochs=1,
                                  constrained_embedding=True, learning_rate=0.1, momentum=0.1, n_sample=14, sample_alpha=0.5,
                                  dropout_p_embed=0.2),
                             dict(sample_store=14 * 30)),
    'bprmax_none_cpustore': (dict(n_items=60, n_events=700, seed=10),       # legacy host-side sample store (gru4rec.py:507-514,551-554,605-615)
                             dict(loss='bpr-max', final_act='elu-0.5', layers=[12], batch_size=6, n_epochs=2, learning_rate=0.1,
                                  momentum=0.2, n_sample=16, sample_alpha=0.75, bpreg=0.5),
                             dict(sample_store=16 * 40, store_type='cpu')),
    'bprmax_randorder_normalinit': (dict(n_items=50, n_events=600, seed=14),   # session order drawn per epoch, no time sort, N(0, sigma) init
                                    dict(loss='bpr-max', final_act='leaky-0.1', hidden_act='selu-1.05-1.67', layers=[10], batch_size=5,
                                         n_epochs=2, learning_rate=0.05, momentum=0.0, n_sample=12, sample_alpha=0.75,
                                         train_random_order=True, time_sort=False, sigma=0.15, init_as_normal=True),
                                    dict(sample_store=12 * 30)),
    'top1max_cpustore_randorder': (dict(n_items=55, n_events=650, seed=15),   # host store with uniform sampling (np.random.choice) + random session order
                                   dict(loss='top1-max', final_act='tanh', layers=[9], batch_size=6, n_epochs=2, learning_rate=0.1,
                                        momentum=0.1, n_sample=12, sample_alpha=0.0, train_random_order=True),
                                   dict(sample_store=12 * 35, store_type='cpu')),
    'xe_none_logq_nosample': (dict(n_items=60, n_events=650, seed=16),        # logQ correction without constrained embedding, in-batch negatives only
                              dict(loss='cross-entropy', final_act='softmax', layers=[10], batch_size=7, n_epochs=1, learning_rate=0.1,
                                   momentum=0.1, n_sample=0, logq=0.7),
                              dict(sample_store=0)),
    'bprmax_shared_3layer_drop': (dict(n_items=60, n_events=800, seed=17),    # paramfiles/retailrocket_bprmax_shared_best.py shape, shrunk
                                  dict(loss='bpr-max', final_act='elu-0.5', layers=[8, 8, 8], batch_size=6, n_epochs=2, learning_rate=0.05,
                                       momentum=0.4, n_sample=20, sample_alpha=0.4, bpreg=1.95, constrained_embedding=True,
                                       dropout_p_embed=0.5, dropout_p_hidden=0.05),
                                  dict(sample_store=20 * 30)),
    'xe_none_cpu_nostore': (dict(n_items=50, n_events=500, seed=18),          # no store: one host draw per mini-batch (gru4rec.py:612-613)
                            dict(loss='cross-entropy', final_act='softmax', layers=[8], batch_size=5, n_epochs=1, learning_rate=0.1,
                                 momentum=0.0, n_sample=9, sample_alpha=0.5),
                            dict(sample_store=0, store_type='cpu')),
    'xe_none_messy_data': (dict(n_items=45, n_events=500, seed=19, item_as_str=True, messy=True),   # single-event sessions, shuffled rows, tied
                           dict(loss='cross-entropy', final_act='softmax', layers=[8], batch_size=6, n_epochs=2, learning_rate=0.1,       # timestamps, string ids
                                momentum=0.0, n_sample=8, sample_alpha=0.75),
                           dict(sample_store=8 * 30)),
    # optimiser variants / clipping / smoothing: implemented by the oracle only (the device path raises NotImplementedError)
    'xe_none_adam': (dict(n_items=50, n_events=500, seed=11),
                     dict(loss='cross-entropy', final_act='softmax', layers=[8], batch_size=5, n_epochs=1, n_sample=10,
                          adapt='adam', adapt_params=[0.9, 0.999], learning_rate=0.01, momentum=0.0),
                     dict(sample_store=10 * 30)),
    'bprmax_none_rmsprop_cap': (dict(n_items=50, n_events=500, seed=12),
                                dict(loss='bpr-max', final_act='elu-0.5', layers=[8], batch_size=5, n_epochs=1, n_sample=10,
                                     adapt='rmsprop', adapt_params=[0.9], learning_rate=0.01, momentum=0.1, grad_cap=0.05),
                                dict(sample_store=10 * 30)),
    'xe_embed_adadelta_smooth': (dict(n_items=50, n_events=500, seed=13),
                                 dict(loss='cross-entropy', final_act='softmax', layers=[8], batch_size=5, n_epochs=1, n_sample=10,
                                      embedding=6, adapt='adadelta', adapt_params=[0.9], learning_rate=1.0, momentum=0.0, smoothing=0.1),
                                 dict(sample_store=10 * 30)),
}


def weights_of(gru):
    w = {}
    for i in range(len(gru.layers)):
        w['Wx%d' % i] = gru.Wx[i].get_value()
        w['Wh%d' % i] = gru.Wh[i].get_value()
        w['Wrz%d' % i] = gru.Wrz[i].get_value()
        w['Bh%d' % i] = gru.Bh[i].get_value()
    w['Wy'] = gru.Wy.get_value()
    w['By'] = gru.By.get_value()
    if getattr(gru, 'embedding', 0) and not gru.constrained_embedding:
        w['E'] = gru.E.get_value()
    return w


def messy(df, seed):
    """Data the reference's preprocessing has to cope with: sessions with a single event (they occupy a lane for zero steps,
    gru4rec.py:596-601,630-636), equal timestamps inside a session, rows in random order."""
    import pandas as pd
    rs = np.random.RandomState(seed + 1000)
    n_sess = int(df['SessionId'].max()) + 1
    extra = pd.DataFrame({'SessionId': (n_sess + np.arange(25)).astype(np.int32),
                          'ItemId': rs.choice(df['ItemId'].unique(), size=25),
                          'Time': rs.uniform(df['Time'].min(), df['Time'].max(), size=25)})
    df = pd.concat([df, extra], ignore_index=True)
    tie = rs.rand(len(df)) < 0.15
    df.loc[tie, 'Time'] = np.floor(df.loc[tie, 'Time'])           # ties: the sort falls back on the input order
    # renumber the sessions by start time so that the train/test split by session id stays a split in time
    first = df.groupby('SessionId')['Time'].min().sort_values(kind='stable')
    remap = pd.Series(np.arange(len(first), dtype=np.int32), index=first.index)
    df['SessionId'] = remap[df['SessionId'].values].values
    return df.iloc[rs.permutation(len(df))].reset_index(drop=True)


def run_one(name, dkw, mkw, fkw):
    dkw = dict(dkw)
    is_messy = dkw.pop('messy', False)
    df = make_sessions(**dkw)
    if is_messy:
        df = messy(df, dkw['seed'])
    train, test = train_test_split(df, 0.25)
    train_in = train.copy()
    gru = ref_gru4rec.GRU4Rec(**mkw)
    theano_shim.LOG_CALLS = True
    del theano_shim.FUNCTION_LOG[:]
    del theano_shim.RANDOM_LOG[:]
    theano_shim.RandomStreams._rid = 0
    ref_gru4rec.mrng = theano_shim.RandomStreams(12345)
    # capture initial weights: init() is called inside fit(); wrap it
    orig_init = gru.init
    init_w = {}

    def init_and_capture(data):
        r = orig_init(data)
        init_w.update(weights_of(gru))
        return r
    gru.init = init_and_capture
    buf = io.StringIO()
    perms = []
    orig_perm = np.random.permutation

    def recording_permutation(x):          # train_random_order (gru4rec.py:593): record the session order of every epoch
        r = orig_perm(x)
        perms.append(np.array(r))
        return r
    np.random.permutation = recording_permutation
    try:
        with contextlib.redirect_stdout(buf):
            gru.fit(train, **fkw)
    finally:
        np.random.permutation = orig_perm
    log = buf.getvalue()
    assert not gru.error_during_train, log
    epoch_loss = [float(x) for x in re.findall(r'loss: ([0-9.]+)', log)]
    calls = list(theano_shim.FUNCTION_LOG)
    rnd = list(theano_shim.RANDOM_LOG)
    theano_shim.LOG_CALLS = False
    out = dict(config_name=name)
    out['train_SessionId'] = train_in['SessionId'].values
    out['train_ItemId'] = train_in['ItemId'].values
    out['train_Time'] = train_in['Time'].values
    out['test_SessionId'] = test['SessionId'].values
    out['test_ItemId'] = test['ItemId'].values
    out['test_Time'] = test['Time'].values
    if mkw.get('train_random_order'):
        out['epoch_orders'] = np.stack(perms)
    out['itemidmap_index'] = gru.itemidmap.index.values
    out['n_items'] = gru.n_items
    for k, v in init_w.items():
        out['init_' + k] = v
    for k, v in weights_of(gru).items():
        out['final_' + k] = v
    for i in range(len(gru.layers)):
        out['final_H%d' % i] = gru.H[i].get_value()
    # train calls: 4 inputs.  generate_samples calls: 0 inputs (ST, STI in `upd`)
    B = mkw['batch_size']
    tr = [c for c in calls if len(c[1]) == 4]
    gen = [(pos, c) for pos, c in enumerate(calls) if len(c[1]) == 0]
    n = len(tr)
    X = np.full((n, B), -1, dtype=np.int64); Y = np.full((n, B), -1, dtype=np.int64)
    R = np.zeros((n, B), dtype=np.int8); M = np.zeros(n, dtype=np.int64); cost = np.zeros(n, dtype=np.float32)
    for s, c in enumerate(tr):
        m = int(c[1][2])
        M[s] = m
        X[s, :m] = c[1][0]; Y[s, :m] = c[1][1][:m]; R[s, :m] = np.asarray(c[1][3]).reshape(-1)
        cost[s] = c[2][0]
    out.update(step_X=X, step_Y=Y, step_R=R, step_M=M, step_cost=cost)
    if fkw.get('store_type', 'gpu') == 'cpu' and mkw['n_sample']:
        # host-side store (gru4rec.py:551-554,605-615): the samples travel in the Y input of every call; regroup them into the
        # stores generate_neg_samples() produced (generate_length rows each, the pointer is not reset between epochs)
        glen = fkw['sample_store'] // mkw['n_sample']
        if glen <= 1:
            glen = 1                                        # no store: generate_neg_samples(pop, 1) per mini-batch
        neg = np.stack([np.asarray(c[1][1][int(c[1][2]):], dtype=np.int64) for c in tr])
        n_st = (n + glen - 1) // glen
        st_all = np.zeros((n_st, glen, mkw['n_sample']), dtype=np.int64)
        for k in range(n):
            st_all[k // glen, k % glen] = neg[k]
        out['sample_stores'] = st_all
        out['store_first_step'] = np.arange(n_st, dtype=np.int64) * glen
        out['store_rows_used'] = np.minimum(glen, n - np.arange(n_st) * glen).astype(np.int64)
        out['host_sampler'] = np.array(1)
    # sample stores, and the index of the first train step served by each store
    stores = []
    first_step = []
    for pos, c in gen:
        st = [u for u in c[3] if u.ndim == 2][0]
        stores.append(st)
        first_step.append(sum(1 for cc in calls[:pos] if len(cc[1]) == 4))
    if stores:
        out['sample_stores'] = np.stack(stores)
        out['store_first_step'] = np.array(first_step, dtype=np.int64)
        out['sample_uniforms'] = np.stack([r[2] for r in rnd if r[1] == 'uniform'])
    out['sampling_P_float32'] = np.zeros(0, dtype=np.float32)
    # dropout masks in call order (rid identifies which dropout site: creation order embed -> hidden layers)
    bins = [r for r in rnd if r[1] == 'binomial']
    rids = sorted(set(r[0] for r in bins))
    for j, rid in enumerate(rids):
        ms = [r[2] for r in bins if r[0] == rid]
        assert len(ms) == n, (len(ms), n)
        width = ms[0].shape[1]
        arr = np.zeros((n, B, width), dtype=np.float32)
        for s, mk in enumerate(ms):
            arr[s, :mk.shape[0]] = mk
        out['dropmask_site%d' % j] = arr
    out['epoch_loss'] = np.array(epoch_loss, dtype=np.float64)
    if name == 'bprmax_none':      # a model pickle written by the REFERENCE class (gru4rec.py:742-767), for the loadmodel compatibility test
        del gru.init                # instance attribute installed above to capture the initial weights (not picklable)
        gru.savemodel(os.path.join(ROOT, 'tests', 'golden', name + '.refmodel.pickle'))
    # evaluation through the reference's evaluate_gpu (batch_size chosen small to exercise lane replacement)
    for mode in ('standard', 'conservative'):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rec, mrr = ref_evaluation.evaluate_gpu(gru, test.copy(), cut_off=[1, 5, 20], batch_size=7, mode=mode)
        out['eval_%s_recall' % mode] = np.array([float(r) for r in rec])
        out['eval_%s_mrr' % mode] = np.array([float(r) for r in mrr])
    # the same with a candidate subset (`items=`, evaluation.py:52-56,84-100): every third item id; targets outside the subset
    # give rank 0 in conservative mode (the reference then reports MRR = inf) -- kept as the reference computes it
    sub_ids = gru.itemidmap.index.values[::3].copy()
    out['eval_items_ids'] = sub_ids
    # ('median' cannot be generated: `others == targets` is a Python bool under Theano's operator overloading, evaluation.py:64)
    for mode in ('standard', 'conservative'):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), np.errstate(divide='ignore', invalid='ignore'):
            rec, mrr = ref_evaluation.evaluate_gpu(gru, test.copy(), items=sub_ids, cut_off=[1, 5, 20], batch_size=7, mode=mode)
        out['eval_items_%s_recall' % mode] = np.array([float(r) for r in rec])
        out['eval_items_%s_mrr' % mode] = np.array([float(r) for r in mrr])
    # predict_next_batch on a fixed probe (reference serving path, gru4rec.py:665-728)
    probe_items = gru.itemidmap.index.values[:5]
    sess = np.arange(5)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        p1 = gru.predict_next_batch(sess, probe_items, None, batch=5)
        p2 = gru.predict_next_batch(sess, probe_items[::-1].copy(), None, batch=5)
    out['predict_probe_items'] = probe_items
    out['predict_out1'] = p1.values
    out['predict_out2'] = p2.values
    # the same probe restricted to a list of items (predict_for_item_ids, gru4rec.py:699-703,719-723): a fresh predict function
    # (the reference compiles it once, for whichever form the first call used)
    gru.predict = None
    sub_pred = sub_ids[:7].copy()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        p3 = gru.predict_next_batch(sess, probe_items, sub_pred, batch=5)
        p4 = gru.predict_next_batch(sess, probe_items[::-1].copy(), sub_pred, batch=5)
    out['predict_sub_items'] = sub_pred
    out['predict_sub_out1'] = p3.values
    out['predict_sub_out2'] = p4.values
    out['model_kwargs'] = np.array(repr(mkw))
    out['fit_kwargs'] = np.array(repr(fkw))
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print('%-28s steps=%d epoch_loss=%s R@20=%.4f -> %s (%d KB)' % (name, n, epoch_loss, out['eval_standard_recall'][2], path, os.path.getsize(path) // 1024))


if __name__ == '__main__':
    sel = sys.argv[1:] or list(CONFIGS)
    for name in sel:
        run_one(name, *CONFIGS[name])
