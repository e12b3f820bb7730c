"""The NumPy oracle replayed against fixtures produced by the reference's own code (oracle/make_golden.py)."""
import numpy as np
import pytest
import gru4rec_oracle as orc
from golden_utils import golden_names, load_golden, frames, init_weights, step_masks, step_samples, fit_data, epoch_order

NAMES = golden_names()


def _model(g):
    return orc.OracleGRU4Rec(**g['model_kwargs'])


@pytest.mark.parametrize('name', NAMES)
def test_init_matches_reference(name):
    g = load_golden(name)
    m = _model(g)
    m.init(int(g['n_items']))
    w = init_weights(g)
    for i in range(len(m.layers)):
        np.testing.assert_array_equal(m.Wx[i], w['Wx'][i])
        np.testing.assert_array_equal(m.Wh[i], w['Wh'][i])
        np.testing.assert_array_equal(m.Wrz[i], w['Wrz'][i])
    np.testing.assert_array_equal(m.Wy, w['Wy'])
    if 'E' in w:
        np.testing.assert_array_equal(m.E, w['E'])


@pytest.mark.parametrize('name', NAMES)
def test_schedule_matches_reference(name):
    g = load_golden(name)
    tr, _ = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    assert list(d['itemids']) == list(g['itemidmap_index'])
    n_sample = mk['n_sample'] if g['fit_kwargs'].get('sample_store', 1) else mk['n_sample']
    n_ep = mk['n_epochs']
    if 'epoch_orders' in g and 'host_sampler' not in g:
        # train_random_order: the orders are np.random.permutation draws on the stream init() seeded (gru4rec.py:254,593)
        m = _model(g)
        m.init(int(g['n_items']))
        for e in range(n_ep):
            np.testing.assert_array_equal(np.random.permutation(len(d['offset_sessions']) - 1), g['epoch_orders'][e])
    k = 0
    for e in range(n_ep):
        steps = orc.build_train_schedule(d['data_items'], d['offset_sessions'], epoch_order(g, d, e), mk['batch_size'], mk['n_sample'])
        for s, st in enumerate(steps):
            k += 1
            M = st['M']
            assert M == g['step_M'][k - 1]
            np.testing.assert_array_equal(st['X'], g['step_X'][k - 1, :M])
            np.testing.assert_array_equal(st['Y'], g['step_Y'][k - 1, :M])
            np.testing.assert_array_equal(st['R'].astype(np.int8), g['step_R'][k - 1, :M])
    assert k == len(g['step_M'])


@pytest.mark.parametrize('name', [n for n in NAMES if 'nosample' not in n])
def test_sample_store_matches_reference(name):
    """sampling CDF (gru4rec.py:543-545,556) + K2 (custom_theano_ops.py:318-349) on the recorded uniforms."""
    g = load_golden(name)
    tr, _ = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    if 'host_sampler' in g:
        # store_type='cpu' (gru4rec.py:507-514): np.searchsorted(pop, np.random.rand(...)) (side='left', float64 CDF) on the global
        # NumPy stream that init() seeded with 42 -- reproduced by seeding, replaying init()'s draws and sampling store by store
        m = _model(g)
        m.init(int(g['n_items']))                          # np.random.seed(42) + the weight draws, as the reference's init()
        alpha = mk.get('sample_alpha', 0.75)
        pop = orc.sampling_cdf(d['supports'], alpha)
        glen = g['sample_stores'].shape[1]
        n_sess = len(d['offset_sessions']) - 1
        # chronological order of the draws: a store when the pointer reaches its end, a session permutation at every epoch start
        events = [(int(fs), 0, k) for k, fs in enumerate(g['store_first_step'])]
        if 'epoch_orders' in g:
            first = 0
            for e in range(mk['n_epochs']):
                events.append((first, 1, e))
                first += len(orc.build_train_schedule(d['data_items'], d['offset_sessions'], g['epoch_orders'][e], mk['batch_size'], mk['n_sample']))
        # at the same step the store created before the epoch loop comes first (step 0); later a permutation (epoch start)
        # precedes the refresh that happens inside the epoch
        events.sort(key=lambda t: (t[0], t[1] if t[0] == 0 else 1 - t[1]))
        for step, kind, k in events:
            if kind == 1:
                np.testing.assert_array_equal(np.random.permutation(n_sess), g['epoch_orders'][k])
                continue
            if alpha:
                st = np.searchsorted(pop, np.random.rand(mk['n_sample'] * glen)).reshape(glen, mk['n_sample'])
            else:
                st = np.random.choice(int(g['n_items']), size=mk['n_sample'] * glen).reshape(glen, mk['n_sample'])
            used = int(g['store_rows_used'][k])
            np.testing.assert_array_equal(st[:used], g['sample_stores'][k][:used])
        return
    P = orc.sampling_cdf(d['supports'], mk.get('sample_alpha', 0.75)).astype(np.float32)
    for k in range(len(g['sample_stores'])):
        st = orc.searchsorted_k2(P, g['sample_uniforms'][k]).reshape(g['sample_stores'][k].shape)
        np.testing.assert_array_equal(st, g['sample_stores'][k])
        st2 = orc.searchsorted_k2_loop(P, g['sample_uniforms'][k][:500])
        np.testing.assert_array_equal(st2, g['sample_stores'][k].reshape(-1)[:500])


@pytest.mark.parametrize('name', NAMES)
def test_training_trajectory_matches_reference(name):
    g = load_golden(name)
    tr, _ = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    m = _model(g)
    m.init(int(g['n_items']))
    if mk.get('logq', 0):
        m.P0 = d['supports'].astype(np.float32)
    costs = []
    k = 0
    bounds = [0]
    for e in range(mk['n_epochs']):
        steps = orc.build_train_schedule(d['data_items'], d['offset_sessions'], epoch_order(g, d, e), mk['batch_size'], mk['n_sample'])
        bounds.append(bounds[-1] + len(steps))
        for h in m.H:
            h[:] = 0
        for st in steps:
            c = m.train_step(st['X'], st['Y'], st['R'], samples=step_samples(g, k), masks=step_masks(g, k, st['M']), slots=st['slots'])
            costs.append(c)
            k += 1
    costs = np.array(costs)
    np.testing.assert_allclose(costs, g['step_cost'], rtol=2e-4, atol=1e-6)
    fw = init_weights(g, 'final_')
    for i in range(len(m.layers)):
        np.testing.assert_allclose(m.Wx[i], fw['Wx'][i], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(m.Wh[i], fw['Wh'][i], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(m.Wrz[i], fw['Wrz'][i], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(m.Bh[i], fw['Bh'][i], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(m.Wy, fw['Wy'], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(m.By, fw['By'], rtol=5e-3, atol=1e-4)
    if 'E' in fw:
        np.testing.assert_allclose(m.E, fw['E'], rtol=5e-3, atol=1e-4)
    # epoch loss as printed by the reference (gru4rec.py:654,661)
    cc = g['step_M'].astype(np.float64)
    n_ep = mk['n_epochs']
    for e in range(n_ep):
        sl = slice(bounds[e], bounds[e + 1])
        avgc = np.sum(costs[sl] * cc[sl]) / np.sum(cc[sl])
        assert abs(avgc - g['epoch_loss'][e]) < 2e-4 * max(1, abs(avgc))


@pytest.mark.parametrize('name', NAMES)
def test_evaluation_matches_reference(name):
    g = load_golden(name)
    tr, te = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    m = _model(g)
    m.set_weights(**init_weights(g, 'final_'))
    m.batch_size = mk['batch_size']
    items, off = orc.prepare_eval_data(te, d['itemidmap'])
    for mode in ('standard', 'conservative'):
        rec, mrr = m.evaluate(items, off, batch_size=7, cut_off=(1, 5, 20), mode=mode)
        np.testing.assert_allclose(rec, g['eval_%s_recall' % mode], rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(mrr, g['eval_%s_mrr' % mode], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize('name', NAMES)
def test_evaluation_with_candidate_items_matches_reference(name):
    """evaluate_gpu(items=...) (evaluation.py:52-56,84-100): ranks against a candidate subset; a target outside the subset has
    conservative rank 0, for which the reference reports MRR = inf -- reproduced as computed."""
    g = load_golden(name)
    tr, te = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    m = _model(g)
    m.set_weights(**init_weights(g, 'final_'))
    m.batch_size = mk['batch_size']
    items, off = orc.prepare_eval_data(te, d['itemidmap'])
    sub = d['itemidmap'][g['eval_items_ids']].values
    n_ev = len(items) - (len(off) - 1)
    for mode in ('standard', 'conservative'):
        rec, mrr = m.evaluate(items, off, batch_size=7, cut_off=(1, 5, 20), mode=mode, items=sub)
        # the final activation is taken over the 4 + |subset| columns only; with softmax two nearly equal probabilities can
        # round to a tie in one exp implementation and not in the other (xe_embed_2layer: one event, rank 7 vs 8), so the
        # bound is ONE rank flip: 1/n in recall, (1/r - 1/(r+1)) / n <= 0.5/n in MRR
        np.testing.assert_allclose(rec, g['eval_items_%s_recall' % mode], rtol=1e-6, atol=1.0 / n_ev + 1e-9)
        np.testing.assert_allclose(mrr, g['eval_items_%s_mrr' % mode], rtol=1e-6, atol=0.5 / n_ev + 1e-9)


@pytest.mark.parametrize('name', NAMES)
def test_predict_next_batch_matches_reference(name):
    """predict_next_batch (gru4rec.py:665-728) on the reference's final weights: all items, and a list of items
    (predict_for_item_ids: the final activation is taken over the listed columns only)."""
    g = load_golden(name)
    tr, _ = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    m = _model(g)
    m.set_weights(**init_weights(g, 'final_'))
    m.batch_size = mk['batch_size']
    probe = d['itemidmap'][g['predict_probe_items']].values
    sub = d['itemidmap'][g['predict_sub_items']].values
    for cols, k1, k2 in ((None, 'predict_out1', 'predict_out2'), (sub, 'predict_sub_out1', 'predict_sub_out2')):
        H = [np.zeros((5, L), dtype=np.float32) for L in mk['layers']]
        y1 = m.predict_step(probe, H, Y=cols)
        y2 = m.predict_step(probe[::-1].copy(), H, Y=cols)
        np.testing.assert_allclose(y1.T, g[k1], rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(y2.T, g[k2], rtol=2e-4, atol=1e-6)
