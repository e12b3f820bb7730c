"""Oracle hand-derived loss/activation gradients vs torch.autograd of a torch restatement of the
reference's forward formulas (gru4rec.py:189-248), in float64."""
import numpy as np
import pytest
import torch
import gru4rec_oracle as orc


def t_act(kind, X):
    k, p1, p2 = kind
    if k == 'linear': return X
    if k == 'relu': return torch.clamp(X, min=0)
    if k == 'tanh': return torch.tanh(X)
    if k == 'leaky': return torch.where(X >= 0, X, p1 * X)
    if k == 'elu': return torch.where(X >= 0, X, p1 * (torch.exp(X) - 1))
    if k == 'selu': return p1 * torch.where(X >= 0, X, p2 * (torch.exp(X) - 1))
    if k == 'softmax':
        e = torch.exp(X - X.max(dim=1, keepdim=True).values)
        return e / e.sum(dim=1, keepdim=True)
    if k == 'softmax_logit':
        Xm = X - X.max(dim=1, keepdim=True).values
        return torch.log(torch.exp(Xm).sum(dim=1, keepdim=True)) - Xm


def t_softmax_neg(X):
    hm = 1.0 - torch.eye(X.shape[0], X.shape[1], dtype=X.dtype)
    X = X * hm
    e = torch.exp(X - X.max(dim=1, keepdim=True).values) * hm
    return e / e.sum(dim=1, keepdim=True)


def t_loss(loss, yhat, M, n_sample, bpreg, smoothing):
    diag = torch.diagonal(yhat)
    if loss == 'cross-entropy':
        if smoothing:
            n_out = M + n_sample
            return torch.sum((1.0 - (n_out / (n_out - 1)) * smoothing) * (-torch.log(diag + 1e-24)) + (smoothing / (n_out - 1)) * torch.sum(-torch.log(yhat + 1e-24), dim=1))
        return torch.sum(-torch.log(diag + 1e-24))
    if loss == 'xe_logit':
        if smoothing:
            n_out = M + n_sample
            return torch.sum((1.0 - (n_out / (n_out - 1)) * smoothing) * diag + (smoothing / (n_out - 1)) * torch.sum(yhat, dim=1))
        return torch.sum(diag)
    if loss == 'bpr':
        return torch.sum(-torch.log(torch.sigmoid(diag[:, None] - yhat)))
    if loss == 'bpr-max':
        s = t_softmax_neg(yhat)
        return torch.sum(-torch.log(torch.sum(torch.sigmoid(diag[:, None] - yhat) * s, dim=1) + 1e-24) + bpreg * torch.sum((yhat ** 2) * s, dim=1))
    if loss == 'top1':
        # as written in the reference (gru4rec.py:242-244): ydiag is a COLUMN, so the difference broadcasts to [M x M]
        ydiag = diag[:, None]
        return torch.sum(torch.mean(torch.sigmoid(-ydiag + yhat) + torch.sigmoid(yhat ** 2), dim=1) - torch.sigmoid(ydiag ** 2) / (M + n_sample))
    if loss == 'top1-max':
        s = t_softmax_neg(yhat)
        return torch.sum(s * (torch.sigmoid(-diag[:, None] + yhat) + torch.sigmoid(yhat ** 2)))


CASES = [
    ('bpr-max', 'elu-0.5', 0.0), ('bpr-max', 'linear', 0.0), ('bpr-max', 'tanh', 0.0), ('bpr-max', 'selu-1.05-1.67', 0.0),
    ('bpr-max', 'leaky-0.1', 0.0), ('bpr-max', 'relu', 0.0),
    ('top1-max', 'tanh', 0.0), ('top1-max', 'elu-1', 0.0), ('bpr', 'linear', 0.0), ('bpr', 'elu-0.5', 0.0),
    ('top1', 'tanh', 0.0), ('cross-entropy', 'softmax', 0.0), ('cross-entropy', 'softmax', 0.1),
    ('xe_logit', 'softmax_logit', 0.0), ('xe_logit', 'softmax_logit', 0.2),
]


@pytest.mark.parametrize('loss,fact,smoothing', CASES)
def test_loss_act_gradient_vs_autograd(loss, fact, smoothing):
    rs = np.random.RandomState(3)
    M, S = 5, 9
    o = rs.randn(M, M + S) * 1.5
    act = orc.parse_act(fact)
    yh = orc.act_fwd(act, o)
    L, dy = orc.loss_and_grad(loss, yh, M, S, bpreg=1.3, smoothing=smoothing)
    do = orc.act_bwd(act, o, yh, dy)
    ot = torch.tensor(o, dtype=torch.float64, requires_grad=True)
    Lt = t_loss(loss, t_act(act, ot), M, S, 1.3, smoothing)
    Lt.backward()
    assert abs(float(Lt.detach()) - float(L)) < 1e-9 * max(1, abs(float(L)))
    np.testing.assert_allclose(do, ot.grad.numpy(), rtol=1e-8, atol=1e-10)


def test_full_step_gradients_finite_difference():
    """float64 central differences of the whole forward (embedding mode, 2 layers, bpr-max)."""
    rs = np.random.RandomState(0)
    m = orc.OracleGRU4Rec(loss='bpr-max', final_act='elu-0.5', layers=[5, 6], batch_size=4, embedding=5, n_sample=6, dtype=np.float64)
    m.init(30)
    m.H = [rs.randn(4, 5) * 0.3, rs.randn(4, 6) * 0.3]
    X = rs.randint(0, 30, 4); Y = rs.randint(0, 30, 4); smp = rs.randint(0, 30, 6); R = np.zeros(4, bool)

    def cost():
        yh, C = m.forward(X, Y, 4, R=R, samples=smp)
        return orc.loss_and_grad(m.loss, yh, 4, m.n_sample, m.bpreg)[0] / m.batch_size, C
    c0, C = cost()
    _, G = m.backward(C, 4)
    eps = 1e-6
    for name, W, dW in [('Wh1', m.Wh[1], G['dWh'][1]), ('Wrz0', m.Wrz[0], G['dWrz'][0]), ('Wx1', m.Wx[1], G['dWx'][1]),
                        ('Wx0', m.Wx[0], G['dWx'][0]), ('Bh0', m.Bh[0], G['dBh'][0])]:
        for _ in range(6):
            idx = tuple(rs.randint(0, s) for s in W.shape)
            old = W[idx]
            W[idx] = old + eps; cp, _ = cost()
            W[idx] = old - eps; cm, _ = cost()
            W[idx] = old
            fd = (cp - cm) / (2 * eps)
            assert abs(fd - dW[idx]) < 1e-7 + 1e-5 * abs(fd), (name, idx, fd, dW[idx])
    # gathered rows: perturb E[X[b]] row b only through the gathered copy -> compare with dSx summed over duplicates
    for b in range(4):
        j = rs.randint(0, 5)
        dup = [bb for bb in range(4) if X[bb] == X[b]]
        old = m.E[X[b], j]
        m.E[X[b], j] = old + eps; cp, _ = cost()
        m.E[X[b], j] = old - eps; cm, _ = cost()
        m.E[X[b], j] = old
        fd = (cp - cm) / (2 * eps)
        assert abs(fd - sum(G['dSx'][bb, j] for bb in dup)) < 1e-7


def test_fp32_trajectory_noise_level():
    """How much of a device-vs-oracle cost difference can be floating-point noise: the SAME oracle trajectory in float32 and in
    float64 stays within ~1e-6 relative in the per-step cost over 60 updates.  The 1e-4 bar on the per-step costs therefore has
    two orders of magnitude of head-room over rounding / summation-order effects; the device tests assert it on every step."""
    from gru4rec_b200.synth import make_sessions
    mk = dict(layers=[24], batch_size=8, n_sample=64, loss='bpr-max', final_act='elu-0.5', learning_rate=0.1, momentum=0.3, sample_alpha=0.0)
    df = make_sessions(n_items=150, n_events=1500, seed=9)
    d = orc.prepare_fit_data(df)
    steps = orc.build_train_schedule(d['data_items'], d['offset_sessions'], d['base_order'], 8, 64)[:60]
    rs = np.random.RandomState(1)
    smp = [rs.randint(0, d['n_items'], 64) for _ in steps]
    runs = []
    for dt in (np.float32, np.float64):
        m = orc.OracleGRU4Rec(dtype=dt, **mk)
        m.init(d['n_items'])
        runs.append(np.array([m.train_step(st['X'], st['Y'], st['R'], samples=smp[k], slots=st['slots']) for k, st in enumerate(steps)], dtype=np.float64))
    rel = np.abs(runs[0] - runs[1]) / np.abs(runs[1])
    assert rel.max() < 1e-5, rel.max()
