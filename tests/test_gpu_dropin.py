"""-m gpu: the drop-in surface end to end -- golden fixtures (made by the reference's own code) replayed through the
CUDA path, GRU4Rec.fit()/evaluate_gpu()/predict_next_batch()/savemodel()/loadmodel() and run.py against the oracle."""
import io
import contextlib
import os
import subprocess
import sys
import numpy as np
import pandas as pd
import pytest
import gru4rec_oracle as orc
from gru4rec_b200 import _lib
from gru4rec_b200.synth import make_sessions, train_test_split
from golden_utils import golden_names, load_golden, frames, init_weights, step_samples, fit_data, epoch_order
from gpu_utils import make_cfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def _oracle_only(mk):
    """training options only the oracle implements: the device path raises NotImplementedError in fit()"""
    return mk.get('adapt', 'adagrad') not in ('adagrad', None) or bool(mk.get('grad_cap', 0)) or bool(mk.get('smoothing', 0))


ORACLE_ONLY = [n for n in golden_names() if _oracle_only(load_golden(n)['model_kwargs'])]
NODROP = [n for n in golden_names() if n not in ORACLE_ONLY and
          not (load_golden(n)['model_kwargs'].get('dropout_p_hidden', 0) or load_golden(n)['model_kwargs'].get('dropout_p_embed', 0))]


@pytest.mark.parametrize('step_mode', [0, 1, 2, 3])
@pytest.mark.parametrize('name', NODROP)
def test_golden_trajectory_through_cuda(name, step_mode):
    """Costs of every mini-batch and the final weights the REFERENCE produced vs. the CUDA path on the same inputs."""
    g = load_golden(name)
    tr, _ = frames(g)
    mk = g['model_kwargs']
    d = fit_data(orc, g, tr)
    S = mk['n_sample']
    rows = g['sample_stores'].shape[1] if 'sample_stores' in g else 0
    eng = _lib.Engine(make_cfg(int(g['n_items']), mk, sample_store=max(rows, 2 if rows else 0) * S, step_mode=step_mode))
    w = init_weights(g)
    for i in range(len(mk['layers'])):
        eng.set('Wx%d' % i, w['Wx'][i]); eng.set('Wh%d' % i, w['Wh'][i]); eng.set('Wrz%d' % i, w['Wrz'][i]); eng.set('Bh%d' % i, w['Bh'][i])
    eng.set('Wy', w['Wy']); eng.set('By', w['By'])
    if 'E' in w:
        eng.set('E', w['E'])
    if mk.get('logq', 0):
        eng.set_logq_support(d['supports'].astype(np.float32))
    costs = []
    k = 0
    for e in range(mk['n_epochs']):
        sched = _lib.Schedule(d['data_items'], d['offset_sessions'], epoch_order(g, d, e), mk['batch_size'], S, mode=0)
        per = sched.n_steps
        eng.reset_hidden()
        done = 0
        while done < per:
            if rows:
                si = int(np.searchsorted(g['store_first_step'], k, side='right') - 1)
                st = g['sample_stores'][si]
                eng.set_sample_store(st if st.shape[0] > 1 else np.vstack([st, st]))    # one draw per mini-batch: a 2-row store
                eng.set_sample_pointer(k - int(g['store_first_step'][si]))
                nxt = int(g['store_first_step'][si + 1]) if si + 1 < len(g['store_first_step']) else 10 ** 9
                n = min(per - done, nxt - k)
            else:
                n = per - done
            costs.append(eng.train_steps(sched, done, n))
            done += n; k += n
    costs = np.concatenate(costs)
    np.testing.assert_allclose(costs, g['step_cost'], rtol=2e-4, atol=1e-6)
    fw = init_weights(g, 'final_')
    np.testing.assert_allclose(eng.get('Wy'), fw['Wy'], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(eng.get('By'), fw['By'], rtol=5e-3, atol=1e-4)
    for i in range(len(mk['layers'])):
        np.testing.assert_allclose(eng.get('Wh%d' % i), fw['Wh'][i], rtol=5e-3, atol=1e-4)
        np.testing.assert_allclose(eng.get('Wx%d' % i), fw['Wx'][i], rtol=5e-3, atol=1e-4)


@pytest.mark.parametrize('name', golden_names())
def test_golden_evaluation_through_cuda(name):
    """evaluate_gpu of the reference (Recall/MRR @1,5,20, two tie modes) and predict_next_batch on the reference's final weights."""
    import gru4rec
    import evaluation
    g = load_golden(name)
    tr, te = frames(g)
    mk = g['model_kwargs']
    gru = gru4rec.GRU4Rec(**mk)
    gru.n_items = int(g['n_items'])
    gru.itemidmap = pd.Series(data=np.arange(gru.n_items), index=g['itemidmap_index'], name='ItemIdx')
    fw = init_weights(g, 'final_')
    host = {'Wy': fw['Wy'], 'By': fw['By']}
    for i in range(len(mk['layers'])):
        host.update({'Wx%d' % i: fw['Wx'][i], 'Wh%d' % i: fw['Wh'][i], 'Wrz%d' % i: fw['Wrz'][i], 'Bh%d' % i: fw['Bh'][i]})
    if 'E' in fw:
        host['E'] = fw['E']
    gru._host = host
    gru.error_during_train = False
    gru.predict = None
    for mode in ('standard', 'conservative'):
        with contextlib.redirect_stdout(io.StringIO()):
            rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), cut_off=[1, 5, 20], batch_size=7, mode=mode)
        np.testing.assert_allclose(rec, g['eval_%s_recall' % mode], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(mrr, g['eval_%s_mrr' % mode], rtol=1e-4, atol=1e-9)
    # candidate subset (evaluate_gpu(items=...)): the reference's own numbers; one rank flip allowed where the reference takes
    # the softmax over the subset columns only (see tests/test_oracle_golden.py)
    n_ev = len(te) - te['SessionId'].nunique()
    for mode in ('standard', 'conservative'):
        with contextlib.redirect_stdout(io.StringIO()):
            rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), items=g['eval_items_ids'], cut_off=[1, 5, 20], batch_size=7, mode=mode)
        np.testing.assert_allclose(rec, g['eval_items_%s_recall' % mode], rtol=1e-4, atol=1.0 / n_ev + 1e-9)
        np.testing.assert_allclose(mrr, g['eval_items_%s_mrr' % mode], rtol=1e-4, atol=0.5 / n_ev + 1e-9)
    with contextlib.redirect_stdout(io.StringIO()):      # the subset is cleared after the call
        rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), cut_off=[1, 5, 20], batch_size=7, mode='standard')
    np.testing.assert_allclose(rec, g['eval_standard_recall'], rtol=1e-4, atol=1e-9)
    probe = g['predict_probe_items']
    p1 = gru.predict_next_batch(np.arange(5), probe, None, batch=5)
    p2 = gru.predict_next_batch(np.arange(5), probe[::-1].copy(), None, batch=5)
    np.testing.assert_allclose(p1.values, g['predict_out1'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(p2.values, g['predict_out2'], rtol=2e-4, atol=1e-6)
    sub = gru.predict_next_batch(np.arange(5) + 100, probe, probe[:3], batch=5)
    assert sub.shape == (3, 5) and list(sub.index) == list(probe[:3])
    # predict_for_item_ids against the reference's own output (fresh sessions -> state reset, as its fresh predict function)
    s1 = gru.predict_next_batch(np.arange(5) + 200, probe, g['predict_sub_items'], batch=5)
    s2 = gru.predict_next_batch(np.arange(5) + 200, probe[::-1].copy(), g['predict_sub_items'], batch=5)
    assert list(s1.index) == list(g['predict_sub_items'])
    np.testing.assert_allclose(s1.values, g['predict_sub_out1'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(s2.values, g['predict_sub_out2'], rtol=2e-4, atol=1e-6)


def _oracle_fit(train, mk, sample_store):
    """What fit() must reproduce: oracle with MRG31k3p sample stores (bit-exact indices) and the same schedule."""
    d = orc.prepare_fit_data(train)
    m = orc.OracleGRU4Rec(**mk)
    m.init(d['n_items'])
    if mk.get('logq', 0):
        m.P0 = d['supports'].astype(np.float32)
    S = mk['n_sample']
    gen_len = sample_store // S
    P = orc.sampling_cdf(d['supports'], mk.get('sample_alpha', 0.75)).astype(np.float32)
    mrg = orc.MRGStreams(12345)
    n = gen_len * S
    st = mrg.substreams(mrg.n_streams(n))
    steps = orc.build_train_schedule(d['data_items'], d['offset_sessions'], d['base_order'], mk['batch_size'], S)
    losses, ptr, store = [], gen_len, None
    for e in range(mk['n_epochs']):
        for h in m.H:
            h[:] = 0
        c, cc = [], []
        for stp in steps:
            if ptr == gen_len:
                store = orc.searchsorted_k2(P, mrg.uniform_from_state(st, n)).reshape(gen_len, S)
                ptr = 0
            c.append(m.train_step(stp['X'], stp['Y'], stp['R'], samples=store[ptr], slots=stp['slots']))
            cc.append(stp['M'])
            ptr += 1
        c, cc = np.array(c), np.array(cc)
        losses.append(np.sum(c * cc) / np.sum(cc))
    return m, d, losses


@pytest.mark.parametrize('mk', [
    dict(loss='bpr-max', final_act='elu-0.5', layers=[24], batch_size=16, n_epochs=2, learning_rate=0.1, momentum=0.3, n_sample=64, sample_alpha=0.0),
    dict(loss='cross-entropy', final_act='softmax', layers=[16], batch_size=8, n_epochs=2, constrained_embedding=True, learning_rate=0.1, momentum=0.2, n_sample=32, sample_alpha=0.5, logq=1.0, dropout_p_hidden=0.3),
])
def test_fit_evaluate_save_load_against_oracle(mk, tmp_path, capsys):
    import gru4rec
    import evaluation
    df = make_sessions(n_items=300, n_events=6000, seed=21, item_as_str=True)
    train, test = train_test_split(df, 0.2)
    store = mk['n_sample'] * 37          # forces several regenerations of the sample store (gru4rec.py:618-621)
    gru = gru4rec.GRU4Rec(**mk)
    gru.fit(train.copy(), sample_store=store)
    out = capsys.readouterr().out
    assert 'Created sample store with 37 batches of samples (type=GPU)' in out
    import re
    dev_losses = [float(x) for x in re.findall(r'Epoch\d+ --> loss: ([0-9.]+)', out)]
    m, d, ref_losses = _oracle_fit(train.copy(), mk, store)
    np.testing.assert_allclose(dev_losses, ref_losses, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(gru.Wy.get_value(), m.Wy, rtol=5e-3, atol=1e-4)
    # evaluation: Recall@20 / MRR@20 within 1e-4 relative of the oracle on the device-trained weights
    m2 = orc.OracleGRU4Rec(**mk)
    host = gru._pull_host()
    nl = len(mk['layers'])
    m2.set_weights(Wx=[host['Wx%d' % i] for i in range(nl)], Wh=[host['Wh%d' % i] for i in range(nl)], Wrz=[host['Wrz%d' % i] for i in range(nl)],
                   Bh=[host['Bh%d' % i] for i in range(nl)], Wy=host['Wy'], By=host['By'])
    items, off = orc.prepare_eval_data(test.copy(), gru.itemidmap)
    r0, q0 = m2.evaluate(items, off, batch_size=50, cut_off=(5, 20), mode='standard')
    rec, mrr = evaluation.evaluate_gpu(gru, test.copy(), cut_off=[5, 20], batch_size=50)
    np.testing.assert_allclose(rec, r0, rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(mrr, q0, rtol=1e-4, atol=1e-9)
    # pickle round trip keeps the scores
    fn = str(tmp_path / 'model.pickle')
    gru.savemodel(fn)
    gru2 = gru4rec.GRU4Rec.loadmodel(fn)
    rec2, mrr2 = evaluation.evaluate_gpu(gru2, test.copy(), cut_off=[5, 20], batch_size=50)
    np.testing.assert_allclose(rec2, rec, rtol=0, atol=1e-12)
    assert isinstance(gru2.Wy, np.ndarray) or hasattr(gru2.Wy, 'get_value')


def test_run_py_cli(tmp_path):
    df = make_sessions(n_items=200, n_events=12000, seed=3)      # run.py evaluates with batch_size=512: needs >= 512 test sessions
    train, test = train_test_split(df, 0.25)
    trp, tep = str(tmp_path / 'train.tsv'), str(tmp_path / 'test.tsv')
    train.to_csv(trp, sep='\t', index=False); test.to_csv(tep, sep='\t', index=False)
    pf = os.path.join(ROOT, 'tests', 'golden', 'params_small.py')
    cmd = [sys.executable, os.path.join(ROOT, 'run.py'), trp, '-pf', pf, '-t', tep, '-m', '5', '20', '-s', str(tmp_path / 'm.pickle'), '-lpm', '-ss', '4096']
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert 'Created sample store with 64 batches of samples (type=GPU)' in out.stdout
    assert 'Epoch2 --> loss:' in out.stdout and 'Recall@20:' in out.stdout and 'PRIMARY METRIC:' in out.stdout
    out2 = subprocess.run([sys.executable, os.path.join(ROOT, 'run.py'), str(tmp_path / 'm.pickle'), '-l', '-t', tep, '-m', '20'], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert out2.returncode == 0, out2.stderr[-2000:]
    r1 = [l for l in out.stdout.splitlines() if l.startswith('Recall@20')][0]
    r2 = [l for l in out2.stdout.splitlines() if l.startswith('Recall@20')][0]
    assert r1 == r2


def test_reference_written_pickle_evaluates_on_device():
    """loadmodel() of a pickle written by the reference class, then evaluate_gpu: the reference's own Recall/MRR."""
    import gru4rec
    import evaluation
    from golden_utils import GOLDEN_DIR
    g = load_golden('bprmax_none')
    _, te = frames(g)
    m = gru4rec.GRU4Rec.loadmodel(os.path.join(GOLDEN_DIR, 'bprmax_none.refmodel.pickle'))
    with contextlib.redirect_stdout(io.StringIO()):
        rec, mrr = evaluation.evaluate_gpu(m, te.copy(), cut_off=[1, 5, 20], batch_size=7, mode='standard')
    np.testing.assert_allclose(rec, g['eval_standard_recall'], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(mrr, g['eval_standard_mrr'], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize('name', ORACLE_ONLY)
def test_unimplemented_training_options_raise_in_fit(name):
    """adam / rmsprop / adadelta, grad_cap and smoothing exist in the oracle only; fit() must fail loudly (no silent fallback),
    while a model the reference trained with them can still be scored (test_golden_evaluation_through_cuda covers that)."""
    import gru4rec
    g = load_golden(name)
    tr, _ = frames(g)
    gru = gru4rec.GRU4Rec(**g['model_kwargs'])
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(NotImplementedError):
            gru.fit(tr.copy(), **g['fit_kwargs'])


@pytest.mark.parametrize('name', [n for n in golden_names() if 'host_sampler' in load_golden(n)])
def test_fit_from_scratch_reproduces_the_reference_run(name):
    """store_type='cpu' makes the whole run a function of NumPy's global stream (seed 42 in init, gru4rec.py:254; samples from
    np.random.rand / np.random.choice, :507-514; session permutations, :593): GRU4Rec.fit() from scratch must reproduce the
    REFERENCE's epoch losses and final weights."""
    import gru4rec
    g = load_golden(name)
    tr, _ = frames(g)
    gru = gru4rec.GRU4Rec(**g['model_kwargs'])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gru.fit(tr.copy(), **g['fit_kwargs'])
    assert not gru.error_during_train
    import re
    losses = [float(x) for x in re.findall(r'loss: ([0-9.]+)', buf.getvalue())]
    np.testing.assert_allclose(losses, g['epoch_loss'], rtol=2e-4, atol=2e-6)
    fw = init_weights(g, 'final_')
    np.testing.assert_allclose(gru.Wy.get_value(), fw['Wy'], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(gru.By.get_value().reshape(-1), fw['By'].reshape(-1), rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(gru.Wh[0].get_value(), fw['Wh'][0], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(gru.Wx[0].get_value(), fw['Wx'][0], rtol=5e-3, atol=1e-4)
