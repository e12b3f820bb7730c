"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: session sharding, agreement on the common step count,
and the merged-update oracle semantics (replicas of the oracle stay identical when every rank applies the merged update)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import gru4rec_oracle as orc
    from gru4rec_b200 import _lib
    from gru4rec_b200.parallel import shard_sessions, common_steps
    from gru4rec_b200.synth import make_sessions
    from gpu_utils import oracle_multi_step
    df = make_sessions(n_items=60, n_events=900, seed=1)
    d = orc.prepare_fit_data(df)
    order = shard_sessions(d['base_order'], rank, world)
    # shards are disjoint and cover everything
    all_orders = [None] * world
    dist.all_gather_object(all_orders, order.tolist())
    flat = sorted(x for o in all_orders for x in o)
    assert flat == sorted(d['base_order'].tolist())
    sched = _lib.Schedule(d['data_items'], d['offset_sessions'], order, 4, 8, mode=0)      # C++ builder on the shard
    n = common_steps(sched.n_steps, dist)
    counts = [None] * world
    dist.all_gather_object(counts, sched.n_steps)
    assert n == min(counts)
    # merged-update semantics on the oracle: each rank applies the merged update -> replicas identical
    mk = dict(layers=[8], batch_size=4, n_sample=8, loss='bpr-max', final_act='elu-0.5', learning_rate=0.1, momentum=0.2)
    m = orc.OracleGRU4Rec(**mk); m.init(d['n_items'])
    e = sched.export()
    Hs = [[np.zeros((4, 8), dtype=np.float32)] for _ in range(world)]
    rs = np.random.RandomState(5)
    for k in range(min(n, 10)):
        mine = dict(X=e['X'][k, :e['M'][k]].astype(np.int64), Y=e['Y'][k, :e['M'][k]].astype(np.int64), R=(e['F'][k, :e['M'][k]] & 1).astype(bool),
                    slots=e['slots'][k, :e['M'][k]].astype(np.int64), samples=np.random.RandomState(50 + rank * 1000 + k).randint(0, d['n_items'], 8))
        inputs = [None] * world
        dist.all_gather_object(inputs, mine)
        oracle_multi_step(m, Hs, inputs)
    w = [None] * world
    dist.all_gather_object(w, m.Wy.tobytes())
    assert all(x == w[0] for x in w)
    q.put((rank, 'ok'))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, 'ok'), (1, 'ok')]
