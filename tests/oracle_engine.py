"""An engine double for CPU tests of the HOST logic of gru4rec_b200.gru4rec.GRU4Rec / evaluation.evaluate_gpu.

`OracleEngine` has the method surface of `gru4rec_b200._lib.Engine` that fit() / evaluate_gpu() / predict_next_batch() use, but
runs every mini-batch through the NumPy oracle (oracle/gru4rec_oracle.py) instead of libg4r.so.  It exists so that the driver's
`-m "not gpu"` suite can exercise the Python class end to end -- sample-store handling, epoch loop, loss weighting, the
multi-process orchestration under gloo -- against the reference's golden runs.  Test infrastructure only: the product never
imports it (the product path has no CPU fallback; `_lib.Engine` raises without a CUDA device)."""
import numpy as np

import gru4rec_oracle as orc
from gpu_utils import oracle_multi_step

_MODE_NAMES = {0: 'standard', 1: 'conservative', 2: 'median', 3: 'tiebreaking'}


def model_kwargs_of(gru):
    """oracle constructor arguments of a GRU4Rec instance"""
    keys = ('loss', 'final_act', 'hidden_act', 'layers', 'n_epochs', 'batch_size', 'dropout_p_hidden', 'dropout_p_embed', 'learning_rate',
            'momentum', 'lmbd', 'embedding', 'n_sample', 'sample_alpha', 'smoothing', 'constrained_embedding', 'adapt', 'adapt_params',
            'grad_cap', 'bpreg', 'logq', 'sigma', 'init_as_normal', 'train_random_order', 'time_sort')
    return {k: getattr(gru, k) for k in keys}


class OracleEngine(object):
    def __init__(self, cfg, mk, device=0):
        self.cfg, self.mk = cfg, dict(mk)
        self.m = orc.OracleGRU4Rec(**mk)
        rng = np.random.get_state()          # the oracle's init() reseeds NumPy's global stream; fit() owns that stream
        self.m.init(int(cfg.n_items))
        np.random.set_state(rng)
        self.n_layers = len(self.m.layers)
        self.S = int(cfg.n_sample)
        self.gen_len = int(cfg.sample_store) // self.S if self.S and int(cfg.sample_store) else 0
        self.store, self.ptr, self.P = None, 0, None
        self.mrg, self.mrg_state = None, None
        self.eval_items, self.He = None, None
        self.dist, self.Hs = None, None
        self.closed = False
        self._exports = {}

    # ---- tensors ----
    def _slot(self, name):
        base, idx = name.rstrip('0123456789'), name[len(name.rstrip('0123456789')):]
        return base, (int(idx) if idx else None)

    def set(self, name, arr):
        base, i = self._slot(name)
        a = np.array(arr, dtype=np.float32)
        if base == 'Bh':
            self.m.Bh[i] = a.reshape(-1)
        elif base in ('Wx', 'Wh', 'Wrz'):
            getattr(self.m, base)[i] = a.reshape(getattr(self.m, base)[i].shape)
        elif base == 'By':
            self.m.By = a.reshape(-1, 1)
        elif base == 'Wy':
            self.m.Wy = a.reshape(self.m.Wy.shape)
        elif base == 'E':
            self.m.E = a.reshape(self.m.E.shape)
        else:
            raise KeyError(name)

    def get(self, name):
        base, i = self._slot(name)
        if base == 'Bh':
            return self.m.Bh[i].reshape(1, -1).copy()
        if base in ('Wx', 'Wh', 'Wrz'):
            return getattr(self.m, base)[i].copy()
        if base == 'By':
            return self.m.By.reshape(-1, 1).copy()
        return getattr(self.m, base).copy()

    # ---- multi-process: replicated oracle, one merged update per lock step (tests/gpu_utils.py::oracle_multi_step) ----
    def init_multi_gpu(self, dist):
        self.dist = dist
        B = self.m.batch_size
        self.Hs = [[np.zeros((B, L), dtype=np.float32) for L in self.m.layers] for _ in range(dist.get_world_size())]

    def sharded(self):
        return False

    def _quiesce(self):
        if self.dist is not None:
            self.dist.barrier()

    # ---- sampling ----
    def set_logq_support(self, P0):
        self.m.P0 = np.asarray(P0, dtype=np.float32)

    def set_sampling_cdf(self, P):
        self.P = np.asarray(P, dtype=np.float32)

    def generate_samples(self):
        n = self.gen_len * self.S
        if self.mrg is None:
            self.mrg = orc.MRGStreams(12345 + int(self.cfg.rank))
            self.mrg_state = self.mrg.substreams(self.mrg.n_streams(n))
        self.store = orc.searchsorted_k2(self.P, self.mrg.uniform_from_state(self.mrg_state, n)).reshape(self.gen_len, self.S)
        self.ptr = 0

    def set_sample_store(self, st):
        self.store = np.asarray(st, dtype=np.int64).reshape(-1, self.S)
        self.ptr = 0

    def sample_store_rows(self):
        return self.gen_len

    def get_sample_pointer(self):
        return self.ptr

    def set_sample_pointer(self, p):
        self.ptr = int(p)

    # ---- training ----
    def reset_hidden(self):
        for h in self.m.H:
            h[:] = 0
        if self.Hs is not None:
            for hs in self.Hs:
                for h in hs:
                    h[:] = 0

    def _export(self, sched):
        key = id(sched)
        if key not in self._exports:
            self._exports = {key: sched.export()}
        return self._exports[key]

    def train_steps(self, sched, first=0, n=None):
        e = self._export(sched)
        n = sched.n_steps - first if n is None else n
        costs = np.empty(n, dtype=np.float32)
        for j in range(n):
            k = first + j
            smp = None
            if self.S:
                if self.store is None or self.ptr >= self.gen_len:
                    assert self.P is not None, 'sample store exhausted and no sampling CDF to refill it from (gru4rec.py:618-621)'
                    self.generate_samples()
                smp = self.store[self.ptr]
                self.ptr += 1
            M = int(e['M'][k])
            mine = dict(X=e['X'][k, :M].astype(np.int64), Y=e['Y'][k, :M].astype(np.int64), R=(e['F'][k, :M] & 1).astype(bool),
                        slots=e['slots'][k, :M].astype(np.int64), samples=smp)
            if self.dist is None:
                costs[j] = self.m.train_step(mine['X'], mine['Y'], mine['R'], samples=smp, slots=mine['slots'])
            else:
                inputs = [None] * self.dist.get_world_size()
                self.dist.all_gather_object(inputs, mine)
                costs[j] = oracle_multi_step(self.m, self.Hs, inputs)[self.dist.get_rank()]
        return costs

    # ---- scoring ----
    def set_eval_items(self, items=None):
        self.eval_items = None if items is None or len(items) == 0 else np.asarray(items, dtype=np.int64)

    def eval_schedule(self, sched, cuts, mode=0):
        m = self.m
        e = sched.export()
        H = [np.zeros((sched.batch_size, L), dtype=np.float32) for L in m.layers]
        rec = np.zeros(len(cuts)); mrr = np.zeros(len(cuts)); n = 0
        for k in range(sched.n_steps):
            M = int(e['M'][k])
            X, Y = e['X'][k, :M].astype(np.int64), e['Y'][k, :M].astype(np.int64)
            ycols = None if self.eval_items is None else np.concatenate([Y, self.eval_items])
            yhat = m.predict_step(X, H, slots=e['slots'][k, :M].astype(np.int64), zero=(e['F'][k, :M] & 2) != 0, Y=ycols)
            rk = m.ranks(yhat, Y, _MODE_NAMES[mode], self.eval_items)
            with np.errstate(divide='ignore', invalid='ignore'):
                for j, c in enumerate(cuts):
                    rec[j] += (rk <= c).sum(); mrr[j] += ((rk <= c) / rk).sum()
            n += M
        return rec, mrr, n

    def reset_eval_hidden(self):
        self.He = None

    def predict(self, X, reset_mask=None):
        if self.He is None or self.He[0].shape[0] != len(X):
            self.He = [np.zeros((len(X), L), dtype=np.float32) for L in self.m.layers]
        return self.m.predict_step(np.asarray(X, dtype=np.int64), self.He, zero=None if reset_mask is None else np.asarray(reset_mask, dtype=bool))

    def close(self):
        self.closed = True


def install(monkeypatch, gru):
    """Route the engines `gru` builds to OracleEngine (for this test only); returns the list of engines created."""
    from gru4rec_b200 import _lib
    made = []

    def make(cfg, device=0):
        eng = OracleEngine(cfg, model_kwargs_of(gru), device)
        made.append(eng)
        return eng
    monkeypatch.setattr(_lib, 'Engine', make)
    return made
