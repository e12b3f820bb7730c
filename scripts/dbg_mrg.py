import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, gru4rec_oracle as orc
from gru4rec_b200 import _lib
from gpu_utils import make_cfg
S, rows, n_items = 64, 40, 300
mk = dict(layers=[8], batch_size=4, n_sample=S, loss='bpr-max', final_act='elu-0.5')
eng = _lib.Engine(make_cfg(n_items, mk, sample_store=S * rows))
n = S * rows
ref = orc.MRGStreams(12345)
st = ref.substreams(ref.n_streams(n))
u = ref.uniform_from_state(st, n)
d = eng.mrg_uniform(n)
bad = np.nonzero(d != u)[0]
print('n_streams', ref.n_streams(n), 'mismatches', len(bad), 'first', bad[:10])
print('dev', d[:8], d[424:430]); print('ref', u[:8], u[424:430])
rs = np.random.RandomState(3)
P = orc.sampling_cdf(rs.randint(1, 100, size=n_items), 0.5).astype(np.float32)
engB = _lib.Engine(make_cfg(n_items, mk, sample_store=S * rows))
engB.set_sampling_cdf(P)
engB.generate_samples()
stB = engB.get_sample_store().reshape(-1)
k2 = orc.searchsorted_k2(P, u)
k2l = orc.searchsorted_k2_loop(P, u)
print('numpy vs loop mismatches', (k2 != k2l).sum())
bad = np.nonzero(stB != k2)[0]
print('store mismatches', len(bad), bad[:20])
ss = eng.searchsorted(P, u)
print('standalone vs numpy', (ss != k2).sum(), 'standalone vs store', (ss != stB).sum())
print(stB[:10], k2[:10])
