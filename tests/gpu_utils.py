"""Shared helpers for the -m gpu parity tests: build an Engine + an oracle with identical state."""
import numpy as np
import gru4rec_oracle as orc
from gru4rec_b200 import _lib


def make_cfg(n_items, mk, sample_store=0, eval_lanes=0, max_resident_steps=0, step_mode=0):
    cfg = _lib.G4RConfig()
    layers = mk.get('layers', [100])
    cfg.n_items = n_items
    cfg.n_layers = len(layers)
    for i, l in enumerate(layers):
        cfg.layers[i] = l
    cfg.batch_size = mk.get('batch_size', 32)
    cfg.constrained_embedding = 1 if mk.get('constrained_embedding') else 0
    cfg.embedding = 0 if mk.get('constrained_embedding') else int(mk.get('embedding', 0) or 0)
    cfg.loss = _lib.LOSS[mk.get('loss', 'bpr-max')]
    cfg.final_act, cfg.final_act_p1, cfg.final_act_p2 = _lib.parse_act(mk.get('final_act', 'linear'))
    cfg.hidden_act, cfg.hidden_act_p1, cfg.hidden_act_p2 = _lib.parse_act(mk.get('hidden_act', 'tanh'))
    cfg.dropout_p_hidden = mk.get('dropout_p_hidden', 0.0)
    cfg.dropout_p_embed = mk.get('dropout_p_embed', 0.0)
    cfg.learning_rate = mk.get('learning_rate', 0.1)
    cfg.momentum = mk.get('momentum', 0.0)
    cfg.lmbd = mk.get('lmbd', 0.0)
    cfg.n_sample = mk.get('n_sample', 2048)
    cfg.sample_alpha = mk.get('sample_alpha', 0.75)
    cfg.smoothing = mk.get('smoothing', 0.0)
    cfg.bpreg = mk.get('bpreg', 1.0)
    cfg.logq = mk.get('logq', 0.0)
    cfg.adapt = _lib.ADAPT[mk.get('adapt', 'adagrad')]
    cfg.sample_store = sample_store
    cfg.dropout_seed = mk.get('dropout_seed', 0)
    cfg.mrg_seed = 12345
    cfg.max_resident_steps = max_resident_steps
    cfg.world_size, cfg.rank = 1, 0
    cfg.eval_batch_size = eval_lanes
    cfg.step_mode = step_mode
    return cfg


def param_names(m):
    names = []
    for i in range(len(m.layers)):
        names += ['Wx%d' % i, 'Wh%d' % i, 'Wrz%d' % i, 'Bh%d' % i]
    names += ['Wy', 'By']
    if m.E is not None:
        names.append('E')
    return names


def oracle_param(m, name):
    if name in ('Wy', 'By', 'E'):
        return getattr(m, name)
    kind, i = name.rstrip('0123456789'), int(name[len(name.rstrip('0123456789')):])
    return getattr(m, kind)[i]


def push_weights(eng, m):
    for name in param_names(m):
        eng.set(name, oracle_param(m, name))
    for i in range(len(m.layers)):
        eng.set('H%d' % i, m.H[i])


def compare_weights(eng, m, rtol, atol, what=''):
    for name in param_names(m):
        dev = eng.get(name)
        ref = np.asarray(oracle_param(m, name)).reshape(dev.shape)
        np.testing.assert_allclose(dev, ref, rtol=rtol, atol=atol, err_msg='%s %s' % (what, name))


def compare_opt_state(eng, m, rtol, atol):
    for (name, slot), val in m.opt.items():
        dev = eng.get('%s.%s' % (name, slot))
        np.testing.assert_allclose(dev, np.asarray(val).reshape(dev.shape), rtol=rtol, atol=atol, err_msg='%s.%s' % (name, slot))


def make_pair(n_items, mk, n_store_rows=0, seed=0, eval_lanes=0, randomize_state=True, step_mode=0):
    """Engine + oracle with identical random weights, hidden state, and (optionally) sample store."""
    rs = np.random.RandomState(seed)
    okw = dict(mk)
    m = orc.OracleGRU4Rec(**okw)
    m.init(n_items)
    if randomize_state:
        for h in m.H:
            h[:] = rs.randn(*h.shape).astype(np.float32) * 0.5
        m.By[:] = rs.randn(*m.By.shape).astype(np.float32) * 0.1
        for b in m.Bh:
            b[:] = rs.randn(*b.shape).astype(np.float32) * 0.1
    S = mk.get('n_sample', 2048)
    cfg = make_cfg(n_items, mk, sample_store=n_store_rows * S, eval_lanes=eval_lanes, step_mode=step_mode)
    eng = _lib.Engine(cfg)
    push_weights(eng, m)
    store = None
    if n_store_rows > 1 and S > 0:
        store = rs.randint(0, n_items, size=(n_store_rows, S)).astype(np.int64)
        # make duplicates likely (inside samples and against targets)
        store[:, :max(1, S // 8)] = rs.randint(0, max(2, n_items // 10), size=(n_store_rows, max(1, S // 8)))
        eng.set_sample_store(store)
    if mk.get('logq', 0):
        P0 = rs.randint(1, 50, size=n_items).astype(np.float32)
        m.P0 = P0
        eng.set_logq_support(P0)
    return eng, m, store, rs
