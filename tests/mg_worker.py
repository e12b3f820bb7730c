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

CASES = {
    'bprmax_none': dict(layers=[24], batch_size=8, n_sample=40, loss='bpr-max', final_act='elu-0.5', learning_rate=0.1, momentum=0.3, sample_alpha=0.0),
    'xe_embed_2layer': dict(layers=[12, 16], batch_size=6, n_sample=30, loss='cross-entropy', final_act='softmax', embedding=12, learning_rate=0.1,
                            dropout_p_hidden=0.2, dropout_p_embed=0.2, lmbd=0.001),
}


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    for name, mk in CASES.items():
        n_items, rows = 90, 400
        B, S = mk['batch_size'], mk['n_sample']
        m = orc.OracleGRU4Rec(**mk)
        m.init(n_items)
        eng = _lib.Engine(_lib.make_config(n_items, mk, sample_store=rows * S, world_size=world, rank=rank), device=local)
        push_weights(eng, m)
        per_rank = []
        for r in range(world):      # every rank reconstructs all ranks' inputs (seeded) to run the merged oracle locally
            df = make_sessions(n_items=n_items, n_events=500, seed=100 + r)
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
        n = min(n, 40)
        eng.set_sample_store(mine['store'])
        eng.init_multi_gpu(dist)
        costs = eng.train_steps(sched, 0, n)
        Hs = [[np.zeros((B, L), dtype=np.float32) for L in mk['layers']] for _ in range(world)]
        ref = []
        for k in range(n):
            inputs = [dict(X=p['steps'][k]['X'], Y=p['steps'][k]['Y'], R=p['steps'][k]['R'], slots=p['steps'][k]['slots'], samples=p['store'][k]) for p in per_rank]
            ref.append(oracle_multi_step(m, Hs, inputs)[rank])
        np.testing.assert_allclose(costs, ref, rtol=3e-4, atol=1e-6)
        compare_weights(eng, m, rtol=3e-3, atol=3e-5, what='%s rank %d' % (name, rank))
        # replicas are bit-identical
        wy = torch.from_numpy(eng.get('Wy')).cuda()
        g = [torch.empty_like(wy) for _ in range(world)]
        dist.all_gather(g, wy)
        assert all(torch.equal(g[0], x) for x in g), 'replicas diverged'
        eng.close()
        if rank == 0:
            print('multi-gpu parity ok:', name, 'steps', n)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
