"""mini-batches/s as the reference reports it: the `Epoch --> loss ... (xx mb/s | yy e/s)` line of GRU4Rec.fit()
(gru4rec.py:654-661: steps of the epoch / wall time of the epoch) on the headline workload, through the user-facing class."""
import os, sys, time, re, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gru4rec
from gru4rec_b200.synth import make_sessions

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
t0 = time.time()
WL = bench.WORKLOADS['cfg2']
df = make_sessions(n_items=WL['n_items'], n_events=n_events, seed=0)
t1 = time.time()
import torch; torch.zeros(1, device='cuda'); torch.cuda.synchronize()     # CUDA context creation is not part of fit()
t1 = time.time()
mk = dict(WL['model']); mk['n_epochs'] = 3
gru = gru4rec.GRU4Rec(**mk)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    gru.fit(df, sample_store=bench.SAMPLE_STORE)
t2 = time.time()
out = buf.getvalue()
print(out.strip())
ep = sum(float(x) for x in re.findall(r'\((\d+\.\d+)s\)', out))
print('synthetic frame: %d events; fit() of %d epochs: %.2f s wall = %.2f s in the epochs + %.2f s before the first epoch (id map, sort check, offsets, weight init, engine, first sample store)' % (len(df), mk['n_epochs'], t2 - t1, ep, t2 - t1 - ep))
