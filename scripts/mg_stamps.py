"""Phase stamps of the row-sharded multi-GPU kernel k_fast_mg (run under torchrun): where the lock step spends its time.
GRU CTA 0: 0 step start | 1 partial dL/dh barrier | 2 b1 done | 3 b2 + group barrier | 4 dense exchange + update | 5 f1 | 6 f2 (h ready)
first helper CTA: 8 dvec ready | 9 input-gradient rows pushed (LL pairs) | 10 owned input rows applied (polling the pairs) | 11 helper barrier |
                  12 next step's owned rows pushed to their requesters | 13 own row received;  first apply CTA: 14 start of apply (after b1) | 15 owned rows applied"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import bench
from gru4rec_b200 import _lib
import gru4rec as g4

rank = int(os.environ['RANK']); local = int(os.environ.get('LOCAL_RANK', rank)); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
wl = bench.WORKLOADS['cfg2']; mk = dict(wl['model'])
cfg = _lib.make_config(wl['n_items'], mk, sample_store=bench.SAMPLE_STORE, max_resident_steps=520, step_mode=2, world_size=world, rank=rank)
eng = _lib.Engine(cfg, device=local)
eng.init_multi_gpu(dist)
gru = g4.GRU4Rec(**mk); gru.n_items = wl['n_items']
for n, w in gru._init_host_weights().items():
    eng.set(n, w)
items, offset, order, supports = bench.build_workload(wl, 1200, seed=rank)
P = np.ones(wl['n_items']).cumsum() / wl['n_items']; P[-1] = 1
eng.set_sampling_cdf(P.astype(np.float32)); eng.generate_samples()
sched = _lib.Schedule(items, offset, order, 32, 2048, mode=0)
eng.reset_hidden()
for k in range(2):
    eng.upload_steps(sched, 256 * k, 256); eng.run_uploaded(256, False)
eng.persistent_stamps(True)
eng.upload_steps(sched, 512, 256); dist.barrier(); c, ms = eng.run_uploaded(256, True)
st = eng.persistent_stamps(False, 256).astype(np.int64)[8:248]
def d(a, b): return float(np.mean(st[:, b] - st[:, a]) / 1000.0)
step = float(np.mean(np.diff(st[:, 0])) / 1000.0)
msg = ('rank %d world %d: lock step %.2f us (events: %.2f) | columns+stats+lossgrad+export (0->1) %.2f | b1 (1->2) %.2f | b2 (2->3) %.2f | dense exchange+update (3->4) %.2f | '
       'f1 (4->5) %.2f | f2 (5->6) %.2f || helper: push (8->9) %.2f | poll+apply (9->10) %.2f | helper barrier (10->11) %.2f | push next rows (11->12) %.2f | receive (12->13) %.2f | '
       'helper chain after dvec (8->13) %.2f vs GRU dense+f1 (3->5) %.2f || apply CTA: start after barrier (1->14) %.2f | poll+apply (14->15) %.2f'
       % (rank, world, step, ms / 256 * 1000, d(0, 1), d(1, 2), d(2, 3), d(3, 4), d(4, 5), d(5, 6), d(8, 9), d(9, 10), d(10, 11), d(11, 12), d(12, 13), d(8, 13), d(3, 5), d(1, 14), d(14, 15)))
for r in range(world):
    if r == rank:
        print(msg, flush=True)
    dist.barrier()
eng.close()
dist.destroy_process_group()
