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
    # ---- row-sharded layout: ownership map, shard sizes, inbox sizing, merged owner plan (host arithmetic of g4r_shard.cuh) ----
    from gru4rec_b200.parallel import owner_of, local_row, shard_rows, sort_columns_owner_major, merged_owner_plan
    lib = _lib.load()
    I = 37483
    rows = [None] * world
    dist.all_gather_object(rows, int(lib.g4r_mg_shard_rows(I, world, rank)))
    assert sum(rows) == I and rows[rank] == shard_rows(I, world, rank)
    for item in (0, 1, 2, 3, I - 1, 12345):
        assert lib.g4r_mg_owner(item, world) == owner_of(item, world)
        assert lib.g4r_mg_local_row(item, world) == local_row(item, world) < rows[owner_of(item, world)]
        assert owner_of(item, world) + world * local_row(item, world) == item
    import ctypes as C
    cfg = _lib.make_config(I, dict(layers=[100], batch_size=32, n_sample=2048, loss='bpr-max', final_act='elu-0.5', momentum=0.3),
                           sample_store=10000000, step_mode=2, world_size=world, rank=rank)
    tot, inbox, inbox_in, dense = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.g4r_mg_segment_bytes(C.byref(cfg), C.byref(tot), C.byref(inbox), C.byref(inbox_in), C.byref(dense)) == 0
    NP, ldW, ld3 = 2080, 104, 300
    assert inbox.value >= 2 * world * NP * ldW * 4          # double-buffered slot [rank][sorted column] per peer: dSy row | dby
    assert inbox_in.value >= 2 * world * 32 * ld3 * 4       # input-row gradients [rank][lane]
    assert dense.value >= 2 * world * 48 * 4 * (3 * 300 + 7)
    assert tot.value >= 3 * rows[rank] * (ldW + ld3) * 4    # parameter + Adagrad + momentum shards of both tables
    # merged owner plan: every (rank, column) of every rank lands in exactly one owner's list, duplicates in (rank, position) order
    rs2 = np.random.RandomState(100 + rank)
    cols = rs2.randint(0, 50, size=40)
    keys, order = sort_columns_owner_major(cols, 50, world)
    allk = [None] * world
    dist.all_gather_object(allk, keys.tolist())
    plan = merged_owner_plan(allk, 50, world, rank)
    counts = [None] * world
    dist.all_gather_object(counts, len(plan))
    assert sum(counts) == 40 * world                          # expected arrivals over all owners == columns scored by all ranks
    assert all(owner_of(it, world) == rank for it, _, _ in plan)
    assert plan == sorted(plan)
    seen = [None] * world
    dist.all_gather_object(seen, [(r, j) for _, r, j in plan])
    flat2 = [x for lst in seen for x in lst]
    assert len(set(flat2)) == len(flat2) == 40 * world
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
