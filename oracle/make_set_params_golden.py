"""TEST INFRASTRUCTURE ONLY.  Runs the reference's GRU4Rec.set_params (gru4rec.py:162-187) through the Theano shim on a list of
keyword sets and records what it printed, the attribute values it left behind and the exception it raised
-> tests/golden/set_params_cases.json.  tests/test_host_logic.py replays the cases through the product class."""
import contextlib, io, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE); sys.path.insert(0, ROOT)
import theano_shim
theano_shim.install()
sys.path.insert(0, '/root/reference')
cwd = os.getcwd()
import gru4rec as ref_gru4rec
os.chdir(cwd)

ATTRS = ['loss', 'final_act', 'hidden_act', 'layers', 'n_epochs', 'batch_size', 'dropout_p_hidden', 'dropout_p_embed', 'learning_rate',
         'momentum', 'lmbd', 'embedding', 'n_sample', 'sample_alpha', 'smoothing', 'constrained_embedding', 'adapt', 'adapt_params',
         'grad_cap', 'bpreg', 'logq', 'sigma', 'init_as_normal', 'train_random_order', 'time_sort', 'session_key', 'item_key', 'time_key']
CASES = [
    dict(loss='bpr-max', layers='100/50', batch_size='64', learning_rate='0.05', constrained_embedding='True', final_act='elu-0.5'),
    dict(embedding='layersize', layers='224'),
    dict(adapt_params='0.9/0.999', adapt='adam', dropout_p_hidden='0.3', train_random_order='1', time_sort='0'),
    dict(n_sample='0', logq='1.0', bpreg='0.5', momentum='0.1', sample_alpha='0.25', n_epochs='3', sigma='0.1', init_as_normal='False',
         grad_cap='1.5', smoothing='0.05', lmbd='0.001', embedding='32', hidden_act='relu', dropout_p_embed='0.1'),
    dict(layers=[100], batch_size=32, learning_rate=0.1, loss='cross-entropy', final_act='softmax', constrained_embedding=True, n_sample=2048),
    dict(loss='top1-max', final_act='selu-1.0507-1.6733', hidden_act='leaky-0.01', session_key='sid', item_key='iid', time_key='ts'),
    dict(foo='1'),
    dict(batch_size='16', constrained_embedding='maybe'),
    dict(loss='hinge'),
    dict(final_act='gelu'),
    dict(hidden_act='softmax'),
    dict(layers='100/x'),
]


def jsonable(v):
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    if isinstance(v, bool):
        return v
    if isinstance(v, (int, float, str)) or v is None:
        return v
    return repr(v)


out = []
for kv in CASES:
    g = ref_gru4rec.GRU4Rec()
    buf = io.StringIO()
    exc = None
    with contextlib.redirect_stdout(buf):
        try:
            g.set_params(**kv)
        except BaseException as e:       # noqa: BLE001
            exc = type(e).__name__
    out.append(dict(kwargs=jsonable(kv), stdout=buf.getvalue(), exception=exc,
                    attrs={a: jsonable(getattr(g, a)) for a in ATTRS}, attr_types={a: type(getattr(g, a)).__name__ for a in ATTRS}))
path = os.path.join(ROOT, 'tests', 'golden', 'set_params_cases.json')
json.dump(out, open(path, 'w'), indent=1)
print('wrote', path, len(out), 'cases;', [c['exception'] for c in out])
