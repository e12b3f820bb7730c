"""Worker of the 2-GPU parity test (launched by tests/test_gpu_multi.py through torch.distributed.run)."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
import gru4rec_oracle as orc
from gru4rec_b200 import _lib
from gru4rec_b200.synth import make_sessions
from gpu_utils import push_weights, compare_weights, oracle_multi_step

# (model keywords, n_items, steps, replicated): the first three run on the row-sharded in-kernel path (k_fast_mg), the
# last two on the replicated NCCL path (shapes the role-specialised kernel does not cover / forced)
CASES = {
    'bprmax_none': (dict(layers=[24], batch_size=8, n_sample=40, loss='bpr-max', final_act='elu-0.5', learning_rate=0.1, momentum=0.3, sample_alpha=0.0), 90, 40, False),
    'xe_none_drop_l2': (dict(layers=[20], batch_size=16, n_sample=64, loss='cross-entropy', final_act='softmax', learning_rate=0.1, dropout_p_hidden=0.2,
                             lmbd=0.001, logq=1.0), 150, 30, False),
    # learning rate as in the single-GPU headline-shape test: with 3000 items every item collects dozens of duplicate updates per
    # lock step, at lr = 0.2 the merged run leaves the stable regime (exp overflow in the oracle) and fp32 noise is amplified
    'headline_shape': (dict(layers=[100], batch_size=32, n_sample=2048, loss='bpr-max', final_act='elu-0.5', learning_rate=0.05, momentum=0.3, sample_alpha=0.0), 3000, 12, False),
    'bprmax_none_replicated': (dict(layers=[24], batch_size=8, n_sample=40, loss='bpr-max', final_act='elu-0.5', learning_rate=0.1, momentum=0.3, sample_alpha=0.0), 90, 20, True),
    'xe_embed_2layer': (dict(layers=[12, 16], batch_size=6, n_sample=30, loss='cross-entropy', final_act='softmax', embedding=12, learning_rate=0.1,
                             dropout_p_hidden=0.2, dropout_p_embed=0.2, lmbd=0.001), 90, 40, False),
}


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    only = os.environ.get('G4R_MG_CASES')
    for name, (mk, n_items, n_max, replicated) in CASES.items():
        if only and name not in only.split(','):
            continue
        rows = 64
        B, S = mk['batch_size'], mk['n_sample']
        if B < world and not replicated and len(mk['layers']) == 1 and not mk.get('embedding'):
            continue
        m = orc.OracleGRU4Rec(**mk)
        m.init(n_items)
        eng = _lib.Engine(_lib.make_config(n_items, mk, sample_store=rows * S, world_size=world, rank=rank, step_mode=2, replicated=replicated), device=local)
        push_weights(eng, m)
        if mk.get('logq', 0):
            P0 = np.random.RandomState(7).randint(1, 50, size=n_items).astype(np.float32)
            m.P0 = P0
            eng.set_logq_support(P0)
        per_rank = []
        for r in range(world):      # every rank reconstructs all ranks' inputs (seeded) to run the merged oracle locally
            df = make_sessions(n_items=n_items, n_events=max(500, 3 * B * (n_max + 8)), seed=100 + r)
            d = orc.prepare_fit_data(df)
            # item ids of the synthetic data are mapped per data set; use raw item indices modulo n_items for a shared vocabulary
            items = (d['data_items'] * 7 + r) % n_items
            steps = orc.build_train_schedule(items, d['offset_sessions'], d['base_order'], B, S)
            store = np.random.RandomState(200 + r).randint(0, n_items, size=(rows, S)).astype(np.int64)
            store[:, :5] = np.random.RandomState(300).randint(0, 6, size=(rows, 5))      # cross-rank duplicates
            per_rank.append(dict(items=items, d=d, steps=steps, store=store))
        mine = per_rank[rank]
        sched = _lib.Schedule(mine['items'], mine['d']['offset_sessions'], mine['d']['base_order'], B, S, mode=0)
        n = min(len(p['steps']) for p in per_rank)
        n = min(n, n_max, rows - 1)
        eng.set_sample_store(mine['store'])
        eng.init_multi_gpu(dist)
        sharded = eng.sharded()
        expect_sharded = (not replicated) and len(mk['layers']) == 1 and not mk.get('embedding')
        assert sharded == expect_sharded, (name, sharded)
        costs = eng.train_steps(sched, 0, n)
        if sharded:
            assert eng.fast_windows()[0] > 0, 'row-sharded path must run the role-specialised kernel' 
        Hs = [[np.zeros((B, L), dtype=np.float32) for L in mk['layers']] for _ in range(world)]
        ref = []
        for k in range(n):
            inputs = [dict(X=p['steps'][k]['X'], Y=p['steps'][k]['Y'], R=p['steps'][k]['R'], slots=p['steps'][k]['slots'], samples=p['store'][k]) for p in per_rank]
            ref.append(oracle_multi_step(m, Hs, inputs)[rank])
        # costs at the north-star tolerance (1e-4 relative); the weights after n steps carry n steps of fp32 accumulation-order noise
        np.testing.assert_allclose(costs, ref, rtol=1e-4, atol=1e-6, err_msg=name)
        compare_weights(eng, m, rtol=3e-3, atol=3e-5, what='%s rank %d' % (name, rank))
        # dense replicas (and, on the replicated path, the tables) are bit-identical on every rank
        for tn in (['Wh0', 'Wrz0', 'Bh0'] + ([] if sharded else ['Wy'])):
            w = torch.from_numpy(eng.get(tn)).cuda()
            g = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(g, w)
            assert all(torch.equal(g[0], x) for x in g), 'replicas diverged: ' + tn
        if sharded:
            # every rank draws its own negatives (own MRG substream block)
            eng.set_sampling_cdf(np.linspace(1.0 / n_items, 1.0, n_items).astype(np.float32))
            eng.generate_samples()
            st = torch.from_numpy(eng.get_sample_store()[:2]).cuda()
            g = [torch.empty_like(st) for _ in range(world)]
            dist.all_gather(g, st)
            assert not torch.equal(g[0], g[1]), 'ranks drew identical negative samples'
        dist.barrier()
        eng.close()
        if rank == 0:
            print('multi-gpu parity ok:', name, 'steps', n, 'sharded' if sharded else 'replicated', 'world', world)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
