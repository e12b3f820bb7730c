"""CPU tests of the HOST side of GRU4Rec.fit() / evaluate_gpu() / predict_next_batch(): the Python class drives an engine double
that computes every mini-batch with the oracle (tests/oracle_engine.py), so what is under test is everything around the step --
id mapping, supports / CDF, sample-store handling and refills, the schedule, the epoch loop and its loss weighting, pickles --
against the REFERENCE's recorded runs (tests/golden) and, for the multi-process orchestration, under gloo with world_size 2."""
import contextlib
import io
import os
import re
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import gru4rec_oracle as orc
from golden_utils import golden_names, load_golden, frames, init_weights
from gru4rec_b200.synth import make_sessions, train_test_split
import oracle_engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SAMPLER = [n for n in golden_names() if 'host_sampler' in load_golden(n)]


def _losses(text):
    return [float(x) for x in re.findall(r'Epoch\d+ --> loss: ([0-9.]+)', text)]


@pytest.mark.parametrize('name', HOST_SAMPLER)
def test_fit_host_logic_reproduces_the_reference_run(name, monkeypatch):
    """store_type='cpu': the whole run is a function of NumPy's global stream (seed 42 in init, gru4rec.py:254; samples from
    np.random, :507-514; session permutations, :593).  The class's own fit() -- with the oracle as the step -- must print the
    REFERENCE's epoch losses and end at its weights: the host logic consumes the stream exactly as the reference does."""
    import gru4rec
    g = load_golden(name)
    tr, te = frames(g)
    gru = gru4rec.GRU4Rec(**g['model_kwargs'])
    made = oracle_engine.install(monkeypatch, gru)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gru.fit(tr.copy(), **g['fit_kwargs'])
    assert not gru.error_during_train and len(made) == 1
    np.testing.assert_allclose(_losses(buf.getvalue()), g['epoch_loss'], rtol=2e-4, atol=2e-6)
    fw = init_weights(g, 'final_')
    np.testing.assert_allclose(gru.Wy.get_value(), fw['Wy'], rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(gru.By.get_value().reshape(-1), fw['By'].reshape(-1), rtol=5e-3, atol=1e-4)
    np.testing.assert_allclose(gru.Wh[0].get_value(), fw['Wh'][0], rtol=5e-3, atol=1e-4)
    # scoring through the class: the reference's own Recall / MRR on its final weights
    import evaluation
    for name_w, w in (('Wy', fw['Wy']), ('By', fw['By'])):
        getattr(gru, name_w).set_value(w)
    nl = len(g['model_kwargs']['layers'])
    for i in range(nl):
        for kind in ('Wx', 'Wh', 'Wrz', 'Bh'):
            getattr(gru, kind)[i].set_value(fw[kind][i])
    if 'E' in fw:
        gru.E.set_value(fw['E'])
    with contextlib.redirect_stdout(io.StringIO()):
        rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), cut_off=[1, 5, 20], batch_size=7, mode='standard')
    np.testing.assert_allclose(rec, g['eval_standard_recall'], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(mrr, g['eval_standard_mrr'], rtol=1e-4, atol=1e-9)


def _oracle_fit(train, mk, sample_store):
    """the literal restatement: oracle + MRG31k3p sample stores + the oracle's own schedule loop"""
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
    dict(loss='bpr-max', final_act='elu-0.5', layers=[12], batch_size=8, n_epochs=2, learning_rate=0.1, momentum=0.3, n_sample=16, sample_alpha=0.0),
    dict(loss='cross-entropy', final_act='softmax', layers=[10], batch_size=6, n_epochs=2, constrained_embedding=True, learning_rate=0.1, momentum=0.2,
         n_sample=12, sample_alpha=0.5, logq=1.0),
])
def test_fit_device_store_orchestration(mk, monkeypatch, tmp_path):
    """store_type='gpu': fit() hands the engine the sampling CDF / logQ supports, lets it refill its store when the pointer wraps
    (gru4rec.py:618-621) and weights the epoch loss by the mini-batch sizes -- same losses as the literal loop; then the pickle
    round trip and predict_next_batch on the class."""
    import gru4rec
    import evaluation
    df = make_sessions(n_items=60, n_events=1200, seed=21, item_as_str=True)
    train, test = train_test_split(df, 0.25)
    store = mk['n_sample'] * 23                                   # several refills per epoch
    gru = gru4rec.GRU4Rec(**mk)
    oracle_engine.install(monkeypatch, gru)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        gru.fit(train.copy(), sample_store=store)
    assert 'Created sample store with 23 batches of samples (type=GPU)' in buf.getvalue()
    m, d, ref_losses = _oracle_fit(train.copy(), mk, store)
    np.testing.assert_allclose(_losses(buf.getvalue()), ref_losses, rtol=1e-6, atol=1e-7)
    np.testing.assert_array_equal(gru.Wy.get_value(), m.Wy)
    with contextlib.redirect_stdout(io.StringIO()):
        rec, mrr = evaluation.evaluate_gpu(gru, test.copy(), cut_off=[5, 20], batch_size=20)
    items, off = orc.prepare_eval_data(test.copy(), gru.itemidmap)
    r0, q0 = m.evaluate(items, off, batch_size=20, cut_off=(5, 20))
    np.testing.assert_allclose(rec, r0, rtol=1e-12); np.testing.assert_allclose(mrr, q0, rtol=1e-12)
    fn = str(tmp_path / 'model.pickle')
    gru.savemodel(fn)
    gru2 = gru4rec.GRU4Rec.loadmodel(fn)
    oracle_engine.install(monkeypatch, gru2)
    with contextlib.redirect_stdout(io.StringIO()):
        rec2, mrr2 = evaluation.evaluate_gpu(gru2, test.copy(), cut_off=[5, 20], batch_size=20)
    assert rec2 == rec and mrr2 == mrr
    ids = gru.itemidmap.index.values[:5]
    p = gru2.predict_next_batch(np.arange(5), ids, None, batch=5)
    assert p.shape == (gru.n_items, 5) and list(p.index) == list(gru.itemidmap.index)
    p2 = gru2.predict_next_batch(np.arange(5), ids[::-1].copy(), ids[:3], batch=5)        # same sessions: the state carries over
    assert p2.shape == (3, 5) and np.isfinite(p2.values).all()


def _job_worker(rank, world, port, tmp, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), G4R_DIST_BACKEND='gloo')
    import torch
    import torch.distributed as dist
    import gru4rec
    import run
    from gru4rec_b200 import _lib
    torch.cuda.current_device = lambda: 0                         # no device in this container; the engine double ignores it
    made = []
    def make(cfg, device=0):
        eng = oracle_engine.OracleEngine(cfg, make.mk, device)
        made.append(eng)
        return eng
    _lib.Engine = make
    ps = 'loss=bpr-max,final_act=elu-0.5,layers=10,batch_size=6,n_sample=12,n_epochs=2,momentum=0.2,learning_rate=0.1,sample_alpha=0.5'
    make.mk = dict(loss='bpr-max', final_act='elu-0.5', layers=[10], batch_size=6, n_sample=12, n_epochs=2, momentum=0.2, learning_rate=0.1, sample_alpha=0.5)
    out = io.StringIO()
    real_stdout = sys.stdout
    sys.stdout = out
    try:
        run.main([os.path.join(tmp, 'train.tsv'), '-ps', ps, '-t', os.path.join(tmp, 'test.tsv'), '-m', '1', '5', '20', '-s', os.path.join(tmp, 'model.pickle'), '-ss', '120'])
    finally:
        text = out.getvalue() if sys.stdout is out else ''        # ranks other than 0 were silenced by run.py
        sys.stdout = real_stdout
    q.put((rank, text, [bool(e.closed) for e in made], [int(e.cfg.world_size) for e in made]))


def test_run_py_job_orchestration_world_size_2_gloo(tmp_path):
    """`torchrun run.py ...` with two processes, on CPU: run.py joins the job, fit() shards the sessions, agrees on the step count,
    prints ONE epoch line for the job, releases the training engines collectively; evaluate_gpu() scores half of the test sessions
    per rank; rank 0 alone prints and saves.  The metrics must equal a single-process evaluation of the saved model."""
    df = make_sessions(n_items=50, n_events=6000, seed=5)             # run.py scores with 512 lanes: >= 512 test sessions
    tr, te = train_test_split(df, 0.3)
    assert te.SessionId.nunique() >= 512
    tr.to_csv(str(tmp_path / 'train.tsv'), sep='\t', index=False); te.to_csv(str(tmp_path / 'test.tsv'), sep='\t', index=False)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_job_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(q.get(timeout=5) for _ in range(2))
    (r0, text0, closed0, worlds0), (r1, text1, closed1, worlds1) = res
    assert (r0, r1) == (0, 1) and text1 == ''                          # rank 1 is silent
    assert worlds0 == worlds1 == [2, 1] and closed0[0] and closed1[0]    # a training engine of the job (released), then a single scoring engine
    epochs = [ln for ln in text0.splitlines() if ln.startswith('Epoch')]
    assert len(epochs) == 2
    losses = _losses(text0)
    assert np.isfinite(losses).all() and losses[1] < losses[0]
    multi = [ln.strip() for ln in text0.splitlines() if ln.startswith('Recall@')]
    assert len(multi) == 3 and os.path.exists(str(tmp_path / 'model.pickle'))
    # one process, the saved model, the whole test set
    import gru4rec
    import evaluation
    from pytest import MonkeyPatch
    mpatch = MonkeyPatch()
    try:
        gru = gru4rec.GRU4Rec.loadmodel(str(tmp_path / 'model.pickle'))
        oracle_engine.install(mpatch, gru)
        import run
        te2 = run.load_data(str(tmp_path / 'test.tsv'), run.build_parser().parse_args(['x']))
        with contextlib.redirect_stdout(io.StringIO()):
            rec, mrr = evaluation.evaluate_gpu(gru, te2, cut_off=[1, 5, 20], batch_size=512)
    finally:
        mpatch.undo()
    single = ['Recall@{}: {:.6f} MRR@{}: {:.6f}'.format(c, rec[i], c, mrr[i]) for i, c in enumerate([1, 5, 20])]
    assert single == multi, (single, multi)
