"""
TEST INFRASTRUCTURE ONLY -- a minimal emulator of the Theano API subset that the reference's
gru4rec.py / evaluation.py / gpu_ops.py use, built on lazy expression nodes evaluated with torch
(CPU, float32) and torch.autograd for T.grad.

Purpose: Theano (third-party, README.md:45 "Theano 1.0.5 or newer") cannot be installed in this
container, so the reference cannot be imported as-is.  With this shim installed into sys.modules
(install()), the UNMODIFIED reference files under /root/reference are imported and their own
model()/loss/RMSprop()/fit()/evaluate_gpu() code builds and runs its graph; oracle/make_golden.py uses
that to produce the fixtures in tests/golden/.  What is emulated (and therefore NOT pinned by the
reference itself): Theano's op semantics (documented behaviour: simultaneous `updates`,
set_subtensor = NumPy fancy assignment (last duplicate wins), inc_subtensor = np.add.at),
MRG_RandomStreams (replaced by a recorded NumPy stream) and the four custom CUDA ops
(custom_theano_ops.py), whose semantics are restated here from the kernel strings.

Never imported by the product path.
"""
import sys
import types
import numpy as np
import torch

floatX = 'float32'
_TORCH_DT = {'float32': torch.float32, 'float64': torch.float64, 'int64': torch.int64, 'int32': torch.int32,
             'int8': torch.int8, 'bool': torch.bool}

RANDOM_LOG = []          # every evaluated random tensor, in evaluation order (cleared by the caller)
FUNCTION_LOG = []        # (function_id, inputs, outputs) per compiled-function call when enabled
LOG_CALLS = False


def _is_var(x):
    return isinstance(x, Var)


def _wrap(x):
    return x if _is_var(x) else Const(x)


def _ev(x, env):
    if _is_var(x):
        return x.eval(env)
    if isinstance(x, slice):
        return slice(_ev_idx(x.start, env), _ev_idx(x.stop, env), _ev_idx(x.step, env))
    if isinstance(x, tuple):
        return tuple(_ev(i, env) for i in x)
    if isinstance(x, list):
        return [_ev(i, env) for i in x]
    return x


def _ev_idx(x, env):
    if x is None:
        return None
    v = _ev(x, env)
    if torch.is_tensor(v) and v.ndim == 0:
        return int(v.item())
    return v


class Var(object):
    def __init__(self, fn=None, inputs=(), name=None):
        self.fn = fn
        self.inputs = list(inputs)
        self.name = name

    def eval(self, env):
        k = id(self)
        if k not in env:
            env[k] = self.fn(*[_ev(i, env) for i in self.inputs])
        return env[k]

    # ---- arithmetic ----
    def _bin(self, other, f, rev=False):
        a, b = (other, self) if rev else (self, other)
        return Var(f, [a, b])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, _mul)
    def __rmul__(self, o): return self._bin(o, _mul, True)
    def __truediv__(self, o): return self._bin(o, _div)
    def __rtruediv__(self, o): return self._bin(o, _div, True)
    def __floordiv__(self, o): return self._bin(o, lambda a, b: a // b)
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b)
    def __rpow__(self, o): return self._bin(o, lambda a, b: a ** b, True)
    def __neg__(self): return Var(lambda a: -a, [self])
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    # NB: like Theano, == is NOT overloaded (variables are dict keys in `updates`)

    def __getitem__(self, idx):
        return Subtensor(self, idx)

    @property
    def T(self):
        return Var(lambda a: a.T, [self])

    @property
    def shape(self):
        return ShapeVar(self)

    def sum(self, axis=None, keepdims=False):
        return tsum(self, axis=axis, keepdims=keepdims)

    def max(self, axis=None, keepdims=False):
        if axis is None:
            return Var(lambda a: a.max(), [self])
        return Var(lambda a: a.max(dim=axis, keepdim=keepdims).values, [self])

    def flatten(self):
        return Var(lambda a: a.reshape(-1), [self])

    def reshape(self, shp):
        return Var(lambda a, s: a.reshape(tuple(int(i) for i in s)), [self, tuple(shp)])


def _mul(a, b):
    if torch.is_tensor(a) and a.dtype in (torch.bool,):
        a = a.to(torch.int8)
    if torch.is_tensor(b) and b.dtype in (torch.bool,):
        b = b.to(torch.int8)
    return a * b


def _div(a, b):
    # Theano: int / int -> float (true division); float32 stays float32
    if torch.is_tensor(a) and not a.dtype.is_floating_point and torch.is_tensor(b) and not b.dtype.is_floating_point:
        return a.to(torch.float64) / b.to(torch.float64)
    if torch.is_tensor(a) and a.dtype == torch.bool:
        a = a.to(torch.float64)
    return a / b


class Const(Var):
    def __init__(self, value):
        Var.__init__(self)
        self.value = value

    def eval(self, env):
        v = self.value
        if isinstance(v, np.ndarray):
            return torch.from_numpy(v)
        return v


class ShapeVar(Var):
    def __init__(self, base):
        Var.__init__(self, lambda a: tuple(a.shape), [base])
        self.base = base

    def __getitem__(self, i):
        return Var(lambda a: int(a.shape[i]), [self.base])

    def __iter__(self):          # `T.eye(*X.shape)` (gru4rec.py:200): only ever used on matrices
        return iter([self[0], self[1]])


class Input(Var):
    def __init__(self, dtype, ndim, name=None):
        Var.__init__(self, name=name)
        self.dtype = dtype
        self.ndim = ndim

    def eval(self, env):
        return env[id(self)]


class Shared(Var):
    def __init__(self, value, name=None, borrow=False):
        Var.__init__(self, name=name)
        self.value = np.array(value, copy=True) if not borrow else np.asarray(value)

    def get_value(self, borrow=False):
        return self.value if borrow else self.value.copy()

    def set_value(self, v, borrow=False):
        self.value = np.asarray(v)

    def eval(self, env):
        k = id(self)
        if k not in env:
            v = np.asarray(self.value)
            t = torch.from_numpy(np.ascontiguousarray(v)).reshape(v.shape)
            if t.dtype.is_floating_point:
                t = t.clone().requires_grad_(True)
            env[k] = t
        return env[k]


class Subtensor(Var):
    def __init__(self, base, idx):
        self.base = base
        self.idx = idx
        Var.__init__(self, None, [base, idx])

    def eval(self, env):
        k = id(self)
        if k not in env:
            env[k] = _index(_ev(self.base, env), _ev(self.idx, env))
        return env[k]


def _index(a, idx):
    def conv(i):
        if torch.is_tensor(i):
            if i.ndim == 0:
                return int(i.item())
            return i.to(torch.int64)
        return i
    if isinstance(idx, tuple):
        return a[tuple(conv(i) for i in idx)]
    return a[conv(idx)]


# ---- theano.tensor functions ----
def tsum(x, axis=None, keepdims=False):
    if isinstance(x, (list, tuple)):
        if len(x) == 0:                                    # as_tensor_variable([]) -> empty vector, whose sum is 0.0
            return Var(lambda: torch.zeros((), dtype=torch.float32), [])
        return Var(lambda *xs: torch.stack([torch.as_tensor(v, dtype=torch.float32) if not torch.is_tensor(v) else v for v in xs]).sum(), list(x))

    def f(a):
        if a.dtype == torch.bool:
            a = a.to(torch.int64)
        if axis is None:
            return a.sum()
        return a.sum(dim=axis, keepdim=keepdims)
    return Var(f, [x])


def tmean(x, axis=None, keepdims=False):
    return Var(lambda a: a.mean() if axis is None else a.mean(dim=axis, keepdim=keepdims), [x])


def _unary(tf):
    return lambda x: Var(tf, [x])


def tdot(a, b):
    return Var(lambda x, y: x @ y, [a, b])


def tswitch(c, a, b):
    def f(cc, aa, bb):
        cc = cc.to(torch.bool) if torch.is_tensor(cc) else bool(cc)
        ref = aa if torch.is_tensor(aa) else bb
        aa = aa if torch.is_tensor(aa) else torch.tensor(aa, dtype=ref.dtype)
        bb = bb if torch.is_tensor(bb) else torch.tensor(bb, dtype=ref.dtype)
        return torch.where(cc, aa, bb)
    return Var(f, [c, a, b])


def tcast(x, dtype):
    def f(a):
        if torch.is_tensor(a):
            return a.to(_TORCH_DT[dtype])
        return torch.tensor(a, dtype=_TORCH_DT[dtype])
    return Var(f, [x])


def teye(n, m=None):
    return Var(lambda a, b: torch.eye(int(a), int(b if b is not None else a), dtype=torch.float32), [n, m])


def tconcatenate(lst, axis=0):
    def f(*xs):
        dts = [x.dtype for x in xs]
        if any(d != dts[0] for d in dts):
            tgt = torch.int64 if not any(d.is_floating_point for d in dts) else torch.float32
            xs = [x.to(tgt) for x in xs]
        return torch.cat(list(xs), dim=axis)
    return Var(f, list(lst))


def tmaximum(a, b):
    def f(x, y):
        if not torch.is_tensor(y):
            y = torch.tensor(y, dtype=x.dtype)
        return torch.maximum(x, y)
    return Var(f, [a, b])


def tdiag(x):
    return Var(lambda a: torch.diagonal(a) if a.ndim == 2 else torch.diag(a), [x])


def _np_detach(t):
    return t.detach().cpu().numpy().copy() if torch.is_tensor(t) else np.asarray(t)


def set_subtensor(sub, val):
    assert isinstance(sub, Subtensor)

    def f(base, idx, v):
        out = _np_detach(base)
        vv = _np_detach(v) if torch.is_tensor(v) else v
        ii = _np_detach(idx) if torch.is_tensor(idx) else idx
        out[ii] = vv                      # NumPy fancy assignment: last duplicate wins (Theano CPU perform())
        return torch.from_numpy(out)
    return Var(f, [sub.base, sub.idx, val])


def inc_subtensor(sub, val):
    assert isinstance(sub, Subtensor)

    def f(base, idx, v):
        out = _np_detach(base)
        vv = _np_detach(v) if torch.is_tensor(v) else v
        ii = _np_detach(idx) if torch.is_tensor(idx) else idx
        if isinstance(ii, np.ndarray):
            np.add.at(out, ii, vv)        # duplicates accumulate (Theano CPU perform())
        else:
            out[ii] += vv
        return torch.from_numpy(out)
    return Var(f, [sub.base, sub.idx, val])


def grad(cost, wrt):
    def f_factory(c, w):
        class G(Var):
            def eval(self, env):
                k = id(self)
                if k not in env:
                    ct = c.eval(env)
                    wt = w.eval(env)
                    g, = torch.autograd.grad(ct, wt, retain_graph=True, allow_unused=True)
                    if g is None:
                        g = torch.zeros_like(wt)
                    env[k] = g.detach()
                return env[k]
        return G()
    if isinstance(wrt, (list, tuple)):
        return [f_factory(cost, w) for w in wrt]
    return f_factory(cost, wrt)


class Function(object):
    _count = 0

    def __init__(self, inputs, outputs=None, updates=None, **kw):
        self.inputs = inputs
        self.outputs = outputs
        self.updates = list(updates.items()) if updates is not None else []
        Function._count += 1
        self.fid = Function._count

    def __call__(self, *args):
        env = {}
        assert len(args) == len(self.inputs), (len(args), len(self.inputs))
        for iv, a in zip(self.inputs, args):
            a = np.asarray(a)
            t = torch.from_numpy(np.ascontiguousarray(a.astype(iv.dtype))).reshape(a.shape)
            if iv.ndim == 0:
                t = t.reshape(())
            env[id(iv)] = t
        outs = self.outputs
        single = not isinstance(outs, (list, tuple))
        out_vars = [] if outs is None else ([outs] if single else list(outs))
        out_vals = [_ev(o, env) for o in out_vars]
        new_vals = [(sv, _ev(_wrap(expr), env)) for sv, expr in self.updates]   # all from OLD values
        for sv, v in new_vals:
            v = _np_detach(v) if torch.is_tensor(v) else np.asarray(v)
            sv.value = np.asarray(v).astype(sv.value.dtype).reshape(sv.value.shape if sv.value.ndim == 0 else v.shape)
        res = [np.asarray(_np_detach(v)) if torch.is_tensor(v) else np.asarray(v) for v in out_vals]
        if LOG_CALLS:
            upd = [np.array(sv.value) for sv, _ in self.updates] if len(self.inputs) == 0 else None
            FUNCTION_LOG.append((self.fid, [np.array(a) for a in args], [r.copy() for r in res], upd))
        if outs is None:
            return None
        return res[0] if single else res


class RandomStreams(object):
    """Stand-in for theano.sandbox.rng_mrg.MRG_RandomStreams; draws come from NumPy and are recorded."""

    _rid = 0

    def __init__(self, seed=12345):
        self.rs = np.random.RandomState(seed)

    @classmethod
    def _next_rid(cls):
        cls._rid += 1
        return cls._rid - 1

    def uniform(self, size=None, low=0.0, high=1.0, dtype=floatX):
        rs = self.rs
        rid = self._next_rid()

        def f(sz):
            shp = tuple(int(s) for s in (sz if isinstance(sz, (tuple, list)) else [sz]))
            u = rs.rand(*shp).astype(np.float32)
            RANDOM_LOG.append((rid, 'uniform', u))
            return torch.from_numpy(u)

        class R(Var):
            def eval(self, env):
                k = id(self)
                if k not in env:
                    env[k] = f(_ev(size, env))
                return env[k]
        return R()

    def binomial(self, size=None, n=1, p=0.5, dtype=floatX):
        rs = self.rs
        rid = self._next_rid()

        def f(sz):
            shp = tuple(int(s) for s in sz)
            u = rs.rand(*shp).astype(np.float32)
            b = (u < np.float32(p)).astype(np.float32)
            RANDOM_LOG.append((rid, 'binomial', b))
            return torch.from_numpy(b)

        class R(Var):
            def eval(self, env):
                k = id(self)
                if k not in env:
                    env[k] = f(_ev(size, env))
                return env[k]
        return R()


# ---- custom ops (custom_theano_ops.py), semantics restated from the kernel strings ----
class GpuExtractDiag2D(object):                      # custom_theano_ops.py:66-78
    def __init__(self, context_name=None, keepdims=False):
        self.keepdims = keepdims

    def __call__(self, x):
        kd = self.keepdims
        return Var(lambda a: torch.diagonal(a).reshape(-1, 1) if kd else torch.diagonal(a), [x])


class GpuBinarySearchSorted(object):                 # custom_theano_ops.py:318-349
    def __init__(self, context_name=None, dtype_int64=False):
        self.dtype_int64 = dtype_int64

    def __call__(self, d, x):
        def f(dd, xx):
            dn = _np_detach(dd); xn = _np_detach(xx)
            ld = len(dn)
            out = np.zeros(len(xn), dtype=np.int64)
            for i, val in enumerate(xn):
                a, b = 0, ld - 1
                if val > dn[b]:
                    a = b = ld
                elif val <= dn[0]:
                    a = b = 0
                while b - a > 0:
                    h = (a + b) // 2
                    if val < dn[h]:
                        b = h
                    else:
                        a = h + 1
                out[i] = b
            return torch.from_numpy(out)
        return Var(f, [d, x])


def install():
    """Insert fake `theano`, `custom_opt`, `custom_theano_ops` modules into sys.modules."""
    th = types.ModuleType('theano')
    th.config = types.SimpleNamespace(floatX=floatX)
    th.shared = lambda value, name=None, borrow=False, **kw: Shared(value, name=name, borrow=borrow)
    th.function = lambda inputs, outputs=None, updates=None, **kw: Function(inputs, outputs, updates, **kw)
    T = types.ModuleType('theano.tensor')
    T.ivector = lambda name=None: Input('int32', 1, name)
    T.iscalar = lambda name=None: Input('int32', 0, name)
    T.bcol = lambda name=None: Input('int8', 2, name)
    T.dot = tdot
    T.exp = _unary(torch.exp)
    T.log = _unary(lambda a: torch.log(a) if torch.is_tensor(a) else float(np.log(a)))
    T.sqrt = _unary(torch.sqrt)
    T.tanh = _unary(torch.tanh)
    T.sum = tsum
    T.mean = tmean
    T.switch = tswitch
    T.ge = lambda a, b: _wrap(a) >= b
    T.gt = lambda a, b: _wrap(a) > b
    T.cast = tcast
    T.eye = teye
    T.concatenate = tconcatenate
    T.maximum = tmaximum
    T.diag = tdiag
    T.grad = lambda cost, wrt, **kw: grad(cost, wrt)
    T.inc_subtensor = inc_subtensor
    T.set_subtensor = set_subtensor
    T.zeros_like = lambda x, dtype=None: Var(lambda a: torch.zeros_like(a, dtype=_TORCH_DT[dtype] if dtype else None), [x])
    T.ones_like = lambda x, dtype=None: Var(lambda a: torch.ones_like(a, dtype=_TORCH_DT[dtype] if dtype else None), [x])
    nnet = types.ModuleType('theano.tensor.nnet')
    nnet.sigmoid = _unary(torch.sigmoid)
    T.nnet = nnet
    th.tensor = T
    th.grad = T.grad
    sandbox = types.ModuleType('theano.sandbox')
    rng_mrg = types.ModuleType('theano.sandbox.rng_mrg')
    rng_mrg.MRG_RandomStreams = RandomStreams
    sandbox.rng_mrg = rng_mrg
    th.sandbox = sandbox
    th.scan = None
    cto = types.ModuleType('custom_theano_ops')
    cto.GpuExtractDiag2D = GpuExtractDiag2D
    cto.GpuBinarySearchSorted = GpuBinarySearchSorted
    copt = types.ModuleType('custom_opt')    # custom_opt.py only re-registers a GPU graph optimizer; no semantics
    sys.modules.update({'theano': th, 'theano.tensor': T, 'theano.tensor.nnet': nnet, 'theano.sandbox': sandbox,
                        'theano.sandbox.rng_mrg': rng_mrg, 'custom_theano_ops': cto, 'custom_opt': copt})
    return th
