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


def _eval_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    os.environ['RANK'] = str(rank); os.environ['WORLD_SIZE'] = str(world); os.environ['LOCAL_RANK'] = str(rank)
    import contextlib
    import io
    import pandas as pd
    import gru4rec_oracle as orc
    from gru4rec_b200 import evaluation, parallel
    from gru4rec_b200.gru4rec import GRU4Rec
    from gru4rec_b200.synth import make_sessions, train_test_split
    assert parallel.init_from_env(backend='gloo') == (world, rank)          # what run.py does under torchrun
    assert parallel.init_from_env(backend='gloo') == (world, rank)          # idempotent
    df = make_sessions(n_items=80, n_events=1500, seed=3)
    tr, te = train_test_split(df, 0.3)
    d = orc.prepare_fit_data(tr)
    mk = dict(layers=[10, 12], batch_size=4, n_sample=8, loss='cross-entropy', final_act='softmax', embedding=9)
    m = orc.OracleGRU4Rec(**mk); m.init(d['n_items'])
    gru = GRU4Rec(**mk)
    gru.n_items, gru.itemidmap, gru.error_during_train = d['n_items'], d['itemidmap'], False
    import oracle_engine
    fake = oracle_engine.OracleEngine(gru._make_config(0, 512, training=False, single=True), mk)      # scores with the oracle (tests/oracle_engine.py)
    fake.m = m
    gru._ensure_engine = lambda lanes: fake
    ti, toff = orc.prepare_eval_data(te, d['itemidmap'])
    n_sess = len(toff) - 1
    # shards: disjoint, complete, balanced
    shards = [None] * world
    dist.all_gather_object(shards, parallel.shard_eval_sessions(n_sess, rank, world).tolist())
    assert sorted(x for s in shards for x in s) == list(range(n_sess)) and max(map(len, shards)) - min(map(len, shards)) <= 1
    cand = d['itemidmap'].index.values[::3]
    assert n_sess >= 100 > len(shards[rank])
    for mode in ('standard', 'conservative', 'median'):
        for items in (None, cand):
            for bs in (7, 100):                   # 100 lanes > the sessions of one shard: the shard runs on fewer lanes
                with contextlib.redirect_stdout(io.StringIO()):
                    rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), items=items, cut_off=[1, 5, 20], batch_size=bs, mode=mode)
                idx = None if items is None else d['itemidmap'][items].values
                rec0, mrr0 = m.evaluate(ti, toff, batch_size=bs, cut_off=(1, 5, 20), mode=mode, items=idx)     # the whole test set on one "device"
                np.testing.assert_allclose(rec, rec0, rtol=1e-12, atol=0)
                np.testing.assert_allclose(mrr, mrr0, rtol=1e-12, atol=0)
                both = [None] * world
                dist.all_gather_object(both, (rec, mrr))
                assert both[0] == both[1]                                     # every rank returns the job's result
    # fewer sessions than lanes: the reference's IndexError on every rank, before anything is scored
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(IndexError):
            evaluation.evaluate_gpu(gru, te.copy(), cut_off=[20], batch_size=n_sess + 1)
    # sharding switched off: every rank scores everything
    os.environ['G4R_EVAL_SHARD'] = '0'
    with contextlib.redirect_stdout(io.StringIO()):
        rec, mrr = evaluation.evaluate_gpu(gru, te.copy(), cut_off=[20], batch_size=7)
    rec0, mrr0 = m.evaluate(ti, toff, batch_size=7, cut_off=(20,))
    np.testing.assert_allclose(rec, rec0, rtol=1e-12); np.testing.assert_allclose(mrr, mrr0, rtol=1e-12)
    # a shared-embedding model has no multi-GPU training path: refused before any engine exists, on every rank
    shared = GRU4Rec(layers=[10], batch_size=4, n_sample=8, loss='cross-entropy', final_act='softmax', constrained_embedding=True)
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(NotImplementedError, match='several GPUs'):
            shared.fit(tr.copy())
    assert shared._engine is None
    # the job-wide epoch line of fit(): sums over ranks
    tot = parallel.allreduce_sum([1.5 + rank, 10.0, 3], dist)
    np.testing.assert_allclose(tot, [1.5 * world + sum(range(world)), 10.0 * world, 3 * world])
    q.put((rank, 'ok'))
    dist.destroy_process_group()


def test_sharded_evaluation_world_size_2_gloo():
    """evaluate_gpu under a 2-process job: every rank scores every second test session, the summed result equals the
    oracle's evaluation of the whole test set (all tie modes, with and without a candidate list)."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29810 + os.getpid() % 150
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, 'ok'), (1, 'ok')]
