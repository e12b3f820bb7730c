"""Shared helpers for the -m gpu parity tests: build an Engine + an oracle with identical state."""
import numpy as np
import gru4rec_oracle as orc
from gru4rec_b200 import _lib


make_cfg = _lib.make_config


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


def oracle_multi_step(m, Hs, inputs):
    """One synchronous data-parallel step of R ranks on the oracle: every rank's forward/backward with the shared
    weights and its own hidden state, then ONE merged update -- row gradients of all ranks concatenated in
    (rank, position) order, dense gradients summed (the stated multi-GPU semantics, gru4rec_b200/csrc/g4r_multi.cuh)."""
    Cs, Gs, costs = [], [], []
    base_seed = m.dropout_seed
    for r, inp in enumerate(inputs):
        M = len(inp['X'])
        m.dropout_seed = (base_seed + r * 0x9E3779B1) & 0xffffffff      # independent dropout masks per rank (g4r_lib.cu layout())
        masks = m.make_masks(M)
        m.dropout_seed = base_seed
        Hr = [h[inp['slots']] for h in Hs[r]]
        X = np.asarray(inp['X'], dtype=np.int64); Y = np.asarray(inp['Y'], dtype=np.int64)
        yhat, C = m.forward(X, Y, M, R=inp['R'], samples=inp.get('samples'), masks=masks, H=Hr)
        cost, G = m.backward(C, M)
        Cs.append(C); Gs.append(G); costs.append(cost)
    nl = len(m.layers)
    Cm = dict(mode=Cs[0]['mode'], X=np.concatenate([C['X'] for C in Cs]), Y=np.concatenate([C['Y'] for C in Cs]),
              Sx=np.vstack([C['Sx'] for C in Cs]), Sy=np.vstack([C['Sy'] for C in Cs]))
    Gm = dict(dSx=np.vstack([G['dSx'] for G in Gs]), dSy=np.vstack([G['dSy'] for G in Gs]), dSBy=np.vstack([G['dSBy'] for G in Gs]))
    for key in ('dWx', 'dWh', 'dWrz', 'dBh'):
        Gm[key] = [None if Gs[0][key][i] is None else sum(G[key][i] for G in Gs) for i in range(nl)]
    m.apply_updates(Cm, Gm, sum(len(i['X']) for i in inputs))
    for r, inp in enumerate(inputs):
        for i in range(nl):
            Hs[r][i][inp['slots']] = Cs[r]['H_new'][i]
    m.step_count += 1
    return costs


def assert_step_costs(costs, ref, err_msg=''):
    """Per-mini-batch costs of a whole trajectory at the north-star tolerance (1e-4 relative, every step).  fp32 rounding alone
    stays two orders of magnitude below it (tests/test_oracle_grads.py::test_fp32_trajectory_noise_level)."""
    costs = np.asarray(costs); ref = np.asarray(ref)
    assert costs.shape == ref.shape, (costs.shape, ref.shape)
    np.testing.assert_allclose(costs, ref, rtol=1e-4, atol=1e-6, err_msg=err_msg)
